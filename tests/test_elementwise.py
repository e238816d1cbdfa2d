"""Memory-bound tokenizer kernels vs torch fp32 references of the same op."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 6e-3)])
@pytest.mark.parametrize("N,H,C,joint", [(3, 16, 512, 1), (2, 32, 128, 1), (4, 16, 256, 2), (1, 64, 768, 1)])
def test_groupnorm_apply(cuda, dtype, tol, N, H, C, joint):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(N, C, H, H, generator=g) * 2 + 0.5).to(dtype)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    pos = torch.randn(joint * H * H, C, generator=g)
    xs = x.float()
    if joint > 1:   # frames of one clip normalised together (kv_norm): stack along H
        xs = xs.view(N // joint, joint, C, H, H).permute(0, 2, 1, 3, 4).reshape(N // joint, C, joint * H, H)
    want = F.silu(F.group_norm(xs, 32, gamma, beta, eps=1e-6))
    xn = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    stats = ops.groupnorm_stats(xn, N // joint, 32, 1e-6)
    y = ops.groupnorm_apply(xn, stats, gamma.to(cuda), beta.to(cuda), True)
    got = y.float().view(N // joint, joint * H, H, C).permute(0, 3, 1, 2)
    assert rel_err(got, want) < tol
    # no SiLU + positional embedding (cross-attention token prep)
    want2 = F.group_norm(xs, 32, gamma, beta, eps=1e-5).permute(0, 2, 3, 1).reshape(N // joint, -1, C) + pos
    stats = ops.groupnorm_stats(xn, N // joint, 32, 1e-5)
    y2 = ops.groupnorm_apply(xn, stats, gamma.to(cuda), beta.to(cuda), False, pos=pos.to(cuda))
    assert rel_err(y2.float().view(N // joint, -1, C), want2) < tol


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 4e-3)])
def test_conv_in_and_out(cuda, dtype, tol):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, T, H, C = 2, 5, 32, 128
    clips = torch.rand(B, T, 3, H, H, generator=g)
    w = torch.randn(C, 3, 3, 3, generator=g) * 0.2
    b = torch.randn(C, generator=g)
    y = ops.conv_in(clips.to(cuda), w.reshape(C, 27).contiguous().to(cuda), b.to(cuda), dtype, 2, 3)   # frames 2..4
    want = F.conv2d(clips[:, 2:5].reshape(-1, 3, H, H), w, b, padding=1)
    assert rel_err(y.float().permute(0, 3, 1, 2), want) < tol
    # decoder head: GN -> SiLU -> conv C->3 scattered into clip slots
    x = torch.randn(B * 3, C, H, H, generator=g).to(dtype)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    w3 = torch.randn(3, C, 3, 3, generator=g) * 0.05
    b3 = torch.randn(3, generator=g)
    want = F.conv2d(F.silu(F.group_norm(x.float(), 32, gamma, beta, eps=1e-6)), w3, b3, padding=1)
    xn = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    stats = ops.groupnorm_stats(xn, B * 3, 32, 1e-6)
    out = torch.zeros(B, T, 3, H, H, device=cuda)
    ops.conv_out3(xn, stats, gamma.to(cuda), beta.to(cuda),
                  w3.permute(0, 2, 3, 1).reshape(3, 9, C).contiguous().to(cuda), b3.to(cuda), out, 2, 3)
    assert rel_err(out[:, 2:5].reshape(-1, 3, H, H), want) < (1e-4 if dtype == torch.float32 else tol)
    assert float(out[:, :2].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_upsample_patchify(cuda, dtype):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 16, 16, 64, generator=g).to(dtype).to(cuda)
    up = ops.upsample2x(x)
    want = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(up.float(), want)
    # patchify: reference expression compressive_vq_model.py:192-195 on NCHW
    d = x.float().permute(0, 3, 1, 2)
    p = 4
    ref = d.permute(0, 2, 3, 1).unfold(1, p, p).unfold(2, p, p).permute(0, 1, 2, 4, 5, 3)
    ref = ref.reshape(ref.shape[0], ref.shape[1] * ref.shape[2], -1)
    got = ops.patchify(x, p)
    assert torch.equal(got.float().view(3, 16, 1024), ref)
    # inverse: reference einsum :248-250
    q = torch.randn(3 * 16, 1024, generator=g).to(dtype).to(cuda)
    r = q.float().reshape(3, 4, 4, p, p, 64)
    r = torch.einsum("nhwpqc->nchpwq", r).reshape(3, 64, 16, 16)
    got = ops.patchify(q, p, inverse=True, frames=3, res=16, ch=64)
    assert torch.equal(got.float().permute(0, 3, 1, 2), r)


def test_token_serialise_roundtrip(cuda):
    from ivideogpt_b200 import ops
    from oracle.vq_model_ref import RefCompressiveVQModel, TINY_CFG
    g = torch.Generator().manual_seed(3)
    B, t, f = 3, 2, 5
    ic = torch.randint(0, 8192, (B, t, 256), generator=g)
    idd = torch.randint(0, 8192, (B, f, 16), generator=g)
    ref = RefCompressiveVQModel(**{**TINY_CFG, "num_vq_embeddings": 8192, "num_dyn_embeddings": 8192})
    want_tok, want_lab = ref.serialise(ic, idd)
    tok, lab = ops.tokens_serialise(ic.to(cuda).contiguous(), idd.to(cuda).contiguous(), B, t, f, 256, 16, 8192, 8192)
    assert torch.equal(tok.cpu(), want_tok) and torch.equal(lab.cpu(), want_lab)
    cb_c = torch.randn(8192, 64, generator=g)
    cb_d = torch.randn(8192, 64, generator=g)
    qc, qd = ops.tokens_gather(tok, cb_c.to(cuda), cb_d.to(cuda), t, f, 256, 16, torch.float32)
    assert torch.equal(qc.cpu(), cb_c[ic.reshape(-1)])
    assert torch.equal(qd.cpu(), cb_d[idd.reshape(-1)])


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Lq,Lmax,pos0", [(2, 3, 130, 136, 0), (1, 2, 64, 72, 0), (2, 2, 70, 200, 65)])
def test_rope_kv_tiled_prefill_kernel_matches_the_scalar_one(cuda, B, H, Lq, Lmax, pos0):
    """The prefill-sized RoPE + cache-append kernel (64-position tiles, vector accesses, V^T through shared memory; bf16) against
    the scalar kernel run in fp32 on the same bf16-exact values: q, K, V^T and the row-major V copy must agree to one bf16
    rounding of the same fp32 arithmetic (partial last tile, odd cache offset included)."""
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(B * 100 + Lq)
    qkv = torch.randn(B * Lq, 3 * H * 64, generator=g).to(torch.bfloat16)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 64, 2).float() / 64))
    fr = torch.arange(Lmax).float()[:, None] * inv[None, :]
    cos, sin = fr.cos().contiguous().to(cuda), fr.sin().contiguous().to(cuda)
    outs = {}
    for dt in (torch.bfloat16, torch.float32):
        q = torch.zeros(B, H, Lq, 64, dtype=dt, device=cuda)
        k = torch.zeros(B, H, Lmax, 64, dtype=dt, device=cuda)
        vt = torch.zeros(B, H, 64, Lmax, dtype=dt, device=cuda)
        vr = torch.zeros(B, H, Lmax, 64, dtype=dt, device=cuda)
        ops.rope_kv(qkv.to(dt).to(cuda), q, k, vt, B, Lq, H, Lmax, pos0, None, cos, sin, v_rows=vr)
        outs[dt] = [t.float().cpu() for t in (q, k, vt, vr)]
    for a, b in zip(outs[torch.bfloat16], outs[torch.float32]):
        want = b.to(torch.bfloat16).float()
        assert float((a - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max()) + 1e-6     # one bf16 ulp at the largest magnitude
        assert float((a != want).float().mean()) < 0.02                                        # fma contraction may flip a last bit
    assert float(outs[torch.bfloat16][2][..., :pos0].abs().max()) == 0.0 if pos0 else True   # nothing written before the offset
