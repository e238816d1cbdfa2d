"""tcgen05 GEMM / implicit-GEMM conv kernels vs fp64 math on identically rounded inputs (tolerances stated)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err, tf32_round, tf32_trunc

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[2, 1], ids=["two-issuers", "one-issuer"], autouse=True)
def mma_issuers(request):
    """Every test of this module runs in both modes of the GEMM / conv kernel: the single bit-reproducible tcgen05.mma issuer
    (default, what the bench runs) and the opt-in second issuing warp on the same accumulator."""
    import torch
    if not torch.cuda.is_available():
        yield
        return
    from ivideogpt_b200 import ops
    ops.set_mma_issuers(request.param)
    try:
        yield
    finally:
        ops.set_mma_issuers(1)

# Tolerances: inputs are pre-rounded to the operand format, so the only error left is fp32 accumulation order
# (+ the kernel's truncation of unrounded fp32 A operands to tf32 where noted).
TOL_EXACT_INPUTS = 2e-5


def _mk(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(*shape, generator=g) * scale
    if dtype == torch.bfloat16:
        return x.to(torch.bfloat16)
    return tf32_round(x)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (256, 256, 256, 0), (100, 72, 192, 64), (64, 2304, 768, 0),
                                      (514, 768, 3072, 128), (300, 16386, 128, 0), (128, 256, 512, 256),
                                      (33, 40, 64, 32)])
def test_gemm_plain(cuda, dtype, M, N, K, bn):
    from ivideogpt_b200 import ops
    a = _mk((M, K), dtype, 1).to(cuda)
    w = _mk((N, K), dtype, 2, 0.1).to(cuda)
    out = ops.gemm(a, w, out_dtype=torch.float32, bn=bn)
    want = a.double() @ w.double().t()
    assert rel_err(out, want) < TOL_EXACT_INPUTS


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_epilogues(cuda, dtype):
    from ivideogpt_b200 import ops
    from ivideogpt_b200._lib import ACT_SILU, ACT_SWIGLU
    M, N, K = 200, 192, 128
    a = _mk((M, K), dtype, 3).to(cuda)
    w = _mk((N, K), dtype, 4, 0.1).to(cuda)
    bias = torch.randn(N, device=cuda)
    res32 = torch.randn(M, N, device=cuda)
    base = a.double() @ w.double().t()
    # bias + fp32 residual, fp32 out
    out = ops.gemm(a, w, bias=bias, residual=res32, out_dtype=torch.float32)
    assert rel_err(out, base + bias.double() + res32.double()) < TOL_EXACT_INPUTS
    # in-place residual accumulate (out aliases residual)
    x = res32.clone()
    ops.gemm(a, w, residual=x, out=x)
    assert rel_err(x, base + res32.double()) < TOL_EXACT_INPUTS
    # SiLU, output in operand dtype (bf16 rounding of the result: 2^-9 relative)
    out = ops.gemm(a, w, bias=bias, act=ACT_SILU)
    assert rel_err(out, F.silu(base + bias.double())) < (4e-3 if dtype == torch.bfloat16 else TOL_EXACT_INPUTS)
    # SwiGLU over interleaved pairs
    out = ops.gemm(a, w, act=ACT_SWIGLU, out_dtype=torch.float32)
    want = F.silu(base[:, 0::2]) * base[:, 1::2]
    assert out.shape == (M, N // 2)
    assert rel_err(out, want) < TOL_EXACT_INPUTS
    # alpha
    out = ops.gemm(a, w, alpha=0.125, out_dtype=torch.float32)
    assert rel_err(out, base * 0.125) < TOL_EXACT_INPUTS


def test_gemm_tf32_truncation_of_unrounded_inputs(cuda):
    """fp32 activations are NOT pre-rounded in the pipeline: the tensor core drops the low 13 mantissa bits."""
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(7)
    a = torch.randn(256, 512, generator=g).to(cuda)
    w = tf32_round(torch.randn(128, 512, generator=g) * 0.1).to(cuda)
    out = ops.gemm(a, w, out_dtype=torch.float32)
    assert rel_err(out, tf32_trunc(a).double() @ w.double().t()) < TOL_EXACT_INPUTS
    assert rel_err(out, a.double() @ w.double().t()) < 1e-3   # vs unrounded math: tf32-level agreement


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_batched_heads(cuda, dtype):
    """The attention-style addressing: per-head K offsets, clip-shared B operand, head-offset output columns."""
    from ivideogpt_b200 import ops
    from ivideogpt_b200._lib import BF16, F32
    code = BF16 if dtype == torch.bfloat16 else F32
    Fr, fpc, heads, Lq, Lk, C = 6, 3, 4, 256, 512, 256
    dh = C // heads
    B = Fr // fpc
    q = _mk((Fr, Lq, C), dtype, 11).to(cuda)
    k = _mk((B, Lk, C), dtype, 12).to(cuda)
    s = torch.empty(Fr * heads, Lq, Lk, dtype=torch.float32, device=cuda)
    ops.gemm_raw(ops.gemm_desc(
        dtype=code, a=q.data_ptr(), lda=C, a_bstride=Lq * C, a_rows=Lq, a_cols=C, a_batches=Fr,
        b=k.data_ptr(), ldb=C, b_bstride=Lk * C, b_rows=Lk, b_cols=C, b_batches=B,
        M=Lq, N=Lk, K=dh, batch=Fr * heads, heads=heads, a_bsel=1, a_bdiv=1, b_bsel=1, b_bdiv=fpc, o_bsel=2,
        a_khead=dh, b_khead=dh, out=s.data_ptr(), ldo=Lk, out_bstride=Lq * Lk, out_dtype=F32, alpha=0.5))
    qh = q.double().view(Fr, Lq, heads, dh).permute(0, 2, 1, 3)
    kh = k.double().view(B, Lk, heads, dh).permute(0, 2, 1, 3).repeat_interleave(fpc, dim=0)
    want = 0.5 * qh @ kh.transpose(-1, -2)
    assert rel_err(s.view(Fr, heads, Lq, Lk), want) < TOL_EXACT_INPUTS
    # P.V with V^T [B, C, Lk]
    p = _mk((Fr * heads, Lq, Lk), dtype, 13, 0.05).to(cuda)
    vt = _mk((B, C, Lk), dtype, 14).to(cuda)
    o = torch.empty(Fr * Lq, C, dtype=torch.float32, device=cuda)
    ops.gemm_raw(ops.gemm_desc(
        dtype=code, a=p.data_ptr(), lda=Lk, a_bstride=Lq * Lk, a_rows=Lq, a_cols=Lk, a_batches=Fr * heads,
        b=vt.data_ptr(), ldb=Lk, b_bstride=C * Lk, b_rows=C, b_cols=Lk, b_batches=B,
        M=Lq, N=dh, K=Lk, batch=Fr * heads, heads=heads, a_bsel=2, b_bsel=1, b_bdiv=fpc, o_bsel=1,
        b_nhead=dh, o_nhead=dh, out=o.data_ptr(), ldo=C, out_bstride=Lq * C, out_dtype=F32))
    vh = vt.double().view(B, heads, dh, Lk).repeat_interleave(fpc, dim=0)           # [Fr, heads, dh, Lk]
    want = (p.double().view(Fr, heads, Lq, Lk) @ vh.transpose(-1, -2)).permute(0, 2, 1, 3).reshape(Fr * Lq, C)
    assert rel_err(o, want) < TOL_EXACT_INPUTS


def _pack_conv(w, dtype, wsc=None):
    p = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    if wsc is not None:
        p = torch.cat([p, wsc.reshape(wsc.shape[0], -1)], dim=1)
    return p.contiguous()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,H,Cin,Cout,stride", [(2, 16, 64, 128, 1), (3, 32, 128, 64, 1), (1, 64, 128, 128, 1),
                                                 (2, 32, 128, 128, 2), (1, 16, 256, 768, 1), (1, 256, 64, 64, 1)])
def test_conv3x3(cuda, dtype, N, H, Cin, Cout, stride):
    from ivideogpt_b200 import ops
    x = _mk((N, Cin, H, H), dtype, 21)
    w = _mk((Cout, Cin, 3, 3), dtype, 22, 0.05)
    b = torch.randn(Cout)
    xin = x.double()
    if stride == 2:
        xin = F.pad(xin, (0, 1, 0, 1))
        want = F.conv2d(xin, w.double(), b.double(), stride=2)
    else:
        want = F.conv2d(xin, w.double(), b.double(), padding=1)
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous().to(cuda), _pack_conv(w, dtype).to(cuda), b.to(cuda),
                      stride=stride, out_dtype=torch.float32)
    assert out.shape == (N, H // stride, H // stride, Cout)
    assert rel_err(out.permute(0, 3, 1, 2), want) < TOL_EXACT_INPUTS


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_conv3x3_fused_shortcut_and_residual(cuda, dtype):
    from ivideogpt_b200 import ops
    N, H, Cin, Cout = 2, 32, 128, 256
    y = _mk((N, Cout, H, H), dtype, 31)          # conv2 input (already normalised+activated)
    x = _mk((N, Cin, H, H), dtype, 32)           # block input feeding the 1x1 shortcut
    w2 = _mk((Cout, Cout, 3, 3), dtype, 33, 0.05)
    wsc = _mk((Cout, Cin, 1, 1), dtype, 34, 0.1)
    b2, bsc = torch.randn(Cout), torch.randn(Cout)
    want = F.conv2d(y.double(), w2.double(), b2.double(), padding=1) + F.conv2d(x.double(), wsc.double(), bsc.double())
    out = ops.conv3x3(y.permute(0, 2, 3, 1).contiguous().to(cuda), _pack_conv(w2, dtype, wsc).to(cuda),
                      (b2 + bsc).to(cuda), x2=x.permute(0, 2, 3, 1).contiguous().to(cuda), out_dtype=torch.float32)
    assert rel_err(out.permute(0, 3, 1, 2), want) < TOL_EXACT_INPUTS
    # identity residual
    r = _mk((N, Cout, H, H), dtype, 35)
    want = F.conv2d(y.double(), w2.double(), b2.double(), padding=1) + r.double()
    out = ops.conv3x3(y.permute(0, 2, 3, 1).contiguous().to(cuda), _pack_conv(w2, dtype).to(cuda), b2.to(cuda),
                      residual=r.permute(0, 2, 3, 1).contiguous().to(cuda), out_dtype=torch.float32)
    assert rel_err(out.permute(0, 3, 1, 2), want) < TOL_EXACT_INPUTS


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("N,H,Cin,Cout,joint", [(4, 16, 64, 768, 2), (2, 32, 128, 128, 1), (3, 64, 128, 256, 1)])
def test_conv3x3_fused_groupnorm_statistics(cuda, dtype, N, H, Cin, Cout, joint):
    """Partial sums emitted by the conv epilogue give the same (mean, rstd) as the stand-alone statistics kernels run
    on the conv output (fp32 output here, so both see identical values)."""
    from ivideogpt_b200 import ops
    x = _mk((N, H, H, Cin), dtype, 41).to(cuda)
    w = _mk((Cout, 9 * Cin), dtype, 42, 0.05).to(cuda)
    b = torch.randn(Cout, device=cuda)
    out = ops.conv3x3(x, w, b, out_dtype=torch.float32, gn_groups=32)
    assert hasattr(out, "gn_part")
    samples = N // joint
    got = ops.groupnorm_stats_from_parts(out, samples, 32, 1e-6)
    want = ops.groupnorm_stats(out, samples, 32, 1e-6)
    assert rel_err(got[..., 0], want[..., 0]) < 1e-4 or float((got[..., 0] - want[..., 0]).abs().max()) < 1e-5
    assert rel_err(got[..., 1], want[..., 1]) < 1e-4


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 8e-3), (torch.float32, 1.5e-3)])
@pytest.mark.parametrize("N,H,W,Cin,Cout,C2,residual", [
    (3, 16, 16, 128, 128, 0, True),       # 128-pixel tile spans 8 rows x 16 columns: every padding side is exercised
    (2, 32, 32, 64, 256, 0, False),       # several tiles per frame, BN = 256
    (2, 16, 16, 128, 64, 128, False),     # fused 1x1 shortcut k-blocks must NOT be normalised
    (5, 8, 8, 256, 64, 0, False),         # two frames per 128-pixel tile is impossible (tw*th = 64 < 128): 8x8 -> tw=8, th=16? guarded below
])
def test_conv_with_fused_input_groupnorm_silu(cuda, dtype, tol, N, H, W, Cin, Cout, C2, residual):
    """conv3x3(silu(GroupNorm(x))) with the normalisation applied to the operand tiles inside the conv kernel (the transform
    warps of gemm_tc.cu) against fp64 math: F.conv2d(F.silu(F.group_norm(x))) on the same rounded x and weights.  The
    tolerance covers the bf16 rounding of the normalised operand + tanh.approx (one rounding more than the two-launch path)."""
    from ivideogpt_b200 import ops
    if H * W < 128:
        pytest.skip("a 128-pixel tile must fit inside one frame")
    G = 32
    x = _mk((N, H, W, Cin), dtype, 11, 1.3) + 0.4
    x = x.to(torch.bfloat16) if dtype == torch.bfloat16 else tf32_round(x)
    w = _mk((Cout, 3, 3, Cin), dtype, 12, 0.05)
    gamma = 1.0 + 0.2 * torch.randn(Cin, generator=torch.Generator().manual_seed(13))
    beta = 0.3 * torch.randn(Cin, generator=torch.Generator().manual_seed(14))
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(15))
    parts = [w.reshape(Cout, 9 * Cin)]
    x2 = w2 = None
    if C2:
        x2 = _mk((N, H, W, C2), dtype, 16)
        w2 = _mk((Cout, C2), dtype, 17, 0.05)
        parts.append(w2)
    wp = torch.cat(parts, dim=1).contiguous().to(cuda)
    res = _mk((N, H, W, Cout), dtype, 18) if residual else None
    xc = x.to(cuda)
    stats = ops.groupnorm_stats(xc, N, G, 1e-6)
    sc, sh = ops.groupnorm_coeff(stats, gamma.to(cuda), beta.to(cuda))
    got = ops.conv3x3(xc, wp, bias.to(cuda), x2=None if x2 is None else x2.to(cuda), residual=None if res is None else res.to(cuda),
                      out_dtype=torch.float32, gn_in=(sc, sh, True))
    xd = x.double().permute(0, 3, 1, 2)
    y = F.silu(F.group_norm(xd, G, gamma.double(), beta.double(), eps=1e-6))
    want = F.conv2d(y, w.double().permute(0, 3, 1, 2), bias.double(), padding=1)
    if C2:
        want = want + F.conv2d(x2.double().permute(0, 3, 1, 2), w2.double()[:, :, None, None])
    if res is not None:
        want = want + res.double().permute(0, 3, 1, 2)
    e = rel_err(got.permute(0, 3, 1, 2), want)
    # and against the two-launch path (normalised copy in HBM, then the plain conv)
    y2 = ops.groupnorm_apply(xc, stats, gamma.to(cuda), beta.to(cuda), True)
    two = ops.conv3x3(y2, wp, bias.to(cuda), x2=None if x2 is None else x2.to(cuda), residual=None if res is None else res.to(cuda),
                      out_dtype=torch.float32)
    e2 = rel_err(two.permute(0, 3, 1, 2), want)
    print(f"\n[fused GN conv {dtype}] rel err fused {e:.3e}, two launches {e2:.3e}")
    assert e < tol, f"fused GroupNorm+SiLU conv rel err {e} (two-launch path: {e2})"


@pytest.fixture
def mh2():
    """The opt-in 256 x 256 CTA tile path (measured slower than the default 128 x 256 tiles, profiles/r02) stays tested."""
    from ivideogpt_b200 import ops
    ops.set_gemm_mh2(True)
    try:
        yield
    finally:
        ops.set_gemm_mh2(False)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_256x256_cta_tiles(cuda, dtype, mh2):
    """Enough 128-row tiles for the 256 x 256 CTA tile path (two M sub-tiles share one weight tile; BN = 256), with an odd
    number of sub-tiles so that the last CTA tile has a phantom second half, bias + residual + SiLU epilogue included."""
    from ivideogpt_b200 import ops
    from ivideogpt_b200._lib import ACT_SILU
    M, N, K = 2 * 300 * 128 + 128 + 57, 512, 192          # 601 full sub-tiles + a ragged one
    a = _mk((M, K), dtype, 21).to(cuda)
    w = _mk((N, K), dtype, 22, 0.1).to(cuda)
    bias = torch.randn(N, device=cuda)
    out = ops.gemm(a, w, bias=bias, out_dtype=torch.float32, bn=256)
    want = a.double() @ w.double().t() + bias.double()
    assert rel_err(out, want) < TOL_EXACT_INPUTS
    res = torch.randn(M, N, device=cuda)
    out2 = ops.gemm(a, w, bias=bias, residual=res, act=ACT_SILU, out_dtype=torch.float32, bn=256)
    assert rel_err(out2, F.silu(want + res.double())) < TOL_EXACT_INPUTS


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("stride,C2", [(1, 0), (1, 128), (2, 0)])
def test_conv_256x256_cta_tiles(cuda, dtype, stride, C2, mh2):
    """The decoder-sized conv launches (hundreds of 128-pixel tiles, Cout = 256) take the 256 x 256 CTA tile path: two spatial
    tiles per CTA (possibly in different frames) against one weight tile; padding, stride 2 and the fused 1x1 shortcut."""
    from ivideogpt_b200 import ops
    N, H, W, Cin, Cout = 37, 64, 64, 128, 256              # 37 * 32 = 1184 tiles at stride 1 (odd count at stride 2: 296 / ...)
    x = _mk((N, H, W, Cin), dtype, 31)
    w = _mk((Cout, 3, 3, Cin), dtype, 32, 0.05)
    bias = torch.randn(Cout, generator=torch.Generator().manual_seed(33))
    parts = [w.reshape(Cout, 9 * Cin)]
    Ho, Wo = H // stride, W // stride
    x2 = w2 = None
    if C2:
        x2 = _mk((N, Ho, Wo, C2), dtype, 34)
        w2 = _mk((Cout, C2), dtype, 35, 0.05)
        parts.append(w2)
    wp = torch.cat(parts, dim=1).contiguous().to(cuda)
    got = ops.conv3x3(x.to(cuda), wp, bias.to(cuda), stride=stride, x2=None if x2 is None else x2.to(cuda), out_dtype=torch.float32)
    xd = x.double().permute(0, 3, 1, 2)
    if stride == 2:
        xd = F.pad(xd, (0, 1, 0, 1))                       # diffusers Downsample2D: pad (0, 1, 0, 1), then stride-2 conv, padding 0
        want = F.conv2d(xd, w.double().permute(0, 3, 1, 2), bias.double(), stride=2)
    else:
        want = F.conv2d(xd, w.double().permute(0, 3, 1, 2), bias.double(), padding=1)
    if C2:
        want = want + F.conv2d(x2.double().permute(0, 3, 1, 2), w2.double()[:, :, None, None])
    assert rel_err(got.permute(0, 3, 1, 2), want) < TOL_EXACT_INPUTS
