"""Tokenizer training (row f3, backward half): every backward op of vq_model/train_plan.py against torch autograd run on the
CPU in float64 on the same inputs.  The forward side uses TF32 tensor cores (fp32 storage), so contractions are compared
at TF32 tolerances; the memory-bound kernels of csrc/tok_train.cu at fp32 tolerances."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

TF32 = 4e-3      # relative error (norm-wise) of a TF32 contraction chain against float64
FP32 = 2e-5


def _graph():
    from ivideogpt_b200.vq_model.plan import TokenizerPlan
    from ivideogpt_b200.vq_model.train_plan import TokenizerTrainGraph
    return TokenizerTrainGraph(None, TokenizerPlan(4, torch.float32))


def _run(graph, out, dout):
    graph.backward_from([(out, dout)])
    return {k: v[1] for k, v in graph.pgrads.items()}


def _nhwc(x):      # [N,C,H,W] -> NHWC contiguous
    return x.permute(0, 2, 3, 1).contiguous()


def test_colsum_reduce_axpby_silu(cuda):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 96, generator=g)
    got = ops.colsum(x.to(cuda))
    assert rel_err(got, x.double().sum(0)) < FP32
    acc = torch.ones(96, device=cuda)
    ops.colsum(x.to(cuda), out=acc, accumulate=True)
    assert rel_err(acc, x.double().sum(0) + 1) < FP32
    y = torch.randn(5, 7, 48, generator=g)
    assert rel_err(ops.reduce_mid(y.to(cuda), 5, 7), y.double().sum(1)) < FP32
    a, b = torch.randn(999, generator=g), torch.randn(999, generator=g)
    assert rel_err(ops.axpby(a.to(cuda), b.to(cuda), 0.5, -2.0), 0.5 * a.double() - 2 * b.double()) < FP32
    xd = a.double().requires_grad_(True)
    F.silu(xd).backward(b.double())
    assert rel_err(ops.silu(a.to(cuda)), F.silu(a.double())) < FP32
    assert rel_err(ops.silu(a.to(cuda), b.to(cuda)), xd.grad) < FP32


@pytest.mark.parametrize("silu,frames_per_sample,with_pos", [(True, 1, False), (False, 1, True), (False, 2, True), (True, 3, False)])
def test_groupnorm_bwd(cuda, silu, frames_per_sample, with_pos):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(1)
    S, H, W, C, G = 3, 8, 8, 64, 4
    N = S * frames_per_sample
    x = torch.randn(N, C, H, W, generator=g) * 2 + 0.5
    norm = nn.GroupNorm(G, C, eps=1e-6)
    norm.weight.data = torch.randn(C, generator=g)
    norm.bias.data = torch.randn(C, generator=g) * 0.3
    pos = nn.Parameter(torch.randn(frames_per_sample * H * W, C, generator=g)) if with_pos else None
    dy = torch.randn(N, C, H, W, generator=g)
    # float64 reference: samples merge consecutive frames (the joint kv_norm of conditional_vae.py:47-50)
    xd = x.double().requires_grad_(True)
    nd = nn.GroupNorm(G, C, eps=1e-6).double()
    nd.weight.data, nd.bias.data = norm.weight.data.double(), norm.bias.data.double()
    xs = xd.view(S, frames_per_sample, C, H, W).permute(0, 2, 1, 3, 4)                  # [S, C, f, H, W]
    y = nd(xs).permute(0, 2, 1, 3, 4).reshape(N, C, H, W)
    if silu:
        y = F.silu(y)
    if with_pos:
        pd = pos.detach().double().requires_grad_(True)
        y = (y.permute(0, 2, 3, 1).reshape(S, frames_per_sample * H * W, C) + pd).reshape(N, H, W, C).permute(0, 3, 1, 2)
    y.backward(dy.double())
    gr = _graph()
    norm, pos = norm.to(cuda), (None if pos is None else nn.Parameter(pos.data.to(cuda)))
    xv = Var(_nhwc(x).to(cuda))
    out = gr.gn(xv, norm, silu, samples=S, pos=pos)
    assert rel_err(out.v, _nhwc(y.detach())) < 1e-4
    pg = _run(gr, out, _nhwc(dy).to(cuda))
    assert rel_err(xv.g, _nhwc(xd.grad)) < 2e-4
    assert rel_err(pg[id(norm.weight)], nd.weight.grad) < 2e-4
    assert rel_err(pg[id(norm.bias)], nd.bias.grad) < 2e-4
    if with_pos:
        assert rel_err(pg[id(pos)], pd.grad) < FP32


@pytest.mark.parametrize("stride,C,k_rows", [(1, 32, None), (2, 32, None), (1, 3, 32), (1, 64, None)])
def test_im2col_t(cuda, stride, C, k_rows):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(2)
    N, H, W = 2, 16, 16
    x = torch.randn(N, C, H, W, generator=g)
    xp = F.pad(x, (0, 1, 0, 1)) if stride == 2 else x
    cols = F.unfold(xp, 3, padding=0 if stride == 2 else 1, stride=stride)            # [N, C*9, L], rows (c, a, b)
    L = cols.shape[-1]
    want = cols.view(N, C, 9, L).permute(2, 1, 0, 3).reshape(9 * C, N * L)            # rows (tap, c), columns (n, pixel)
    got = ops.im2col3x3_t(_nhwc(x).to(cuda), stride, k_rows=k_rows).cpu()
    assert torch.equal(got[: 9 * C], want)
    if k_rows:
        assert not got[9 * C:].any()


@pytest.mark.parametrize("case", ["plain", "stride2", "shortcut", "residual"])
def test_conv_bwd(cuda, case):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(3)
    N, H, W, Ci, Co = 4, 32, 32, 64, 128
    if case == "residual":
        Co = Ci
    stride = 2 if case == "stride2" else 1
    conv = nn.Conv2d(Ci, Co, 3, stride=stride, padding=0 if stride == 2 else 1)
    x = torch.randn(N, Ci, H, W, generator=g)
    C2 = 32
    sc = nn.Conv2d(C2, Co, 1) if case == "shortcut" else None
    x2 = torch.randn(N, C2, H, W, generator=g) if sc is not None else None
    res = torch.randn(N, Co, H, W, generator=g) if case == "residual" else None
    dy = torch.randn(N, Co, H // stride, W // stride, generator=g)
    cd = nn.Conv2d(Ci, Co, 3, stride=stride, padding=0 if stride == 2 else 1).double()
    cd.load_state_dict({k: v.double() for k, v in conv.state_dict().items()})
    xd = x.double().requires_grad_(True)
    y = cd(F.pad(xd, (0, 1, 0, 1)) if stride == 2 else xd)
    if sc is not None:
        sd = nn.Conv2d(C2, Co, 1).double()
        sd.load_state_dict({k: v.double() for k, v in sc.state_dict().items()})
        x2d = x2.double().requires_grad_(True)
        y = y + sd(x2d)
    if res is not None:
        rd = res.double().requires_grad_(True)
        y = y + rd
    y.backward(dy.double())
    gr = _graph()
    conv = conv.to(cuda)
    sc = None if sc is None else sc.to(cuda)
    xv = Var(_nhwc(x).to(cuda))
    x2v = None if x2 is None else Var(_nhwc(x2).to(cuda))
    rv = None if res is None else Var(_nhwc(res).to(cuda))
    out = gr.conv(xv, conv, stride=stride, shortcut=sc, x2=x2v, residual=rv)
    assert rel_err(out.v, _nhwc(y.detach())) < TF32
    pg = _run(gr, out, _nhwc(dy).to(cuda))
    assert rel_err(xv.g, _nhwc(xd.grad)) < TF32
    assert rel_err(pg[id(conv.weight)], cd.weight.grad) < TF32
    assert rel_err(pg[id(conv.bias)], cd.bias.grad) < FP32 * 10
    if sc is not None:
        assert rel_err(x2v.g, _nhwc(x2d.grad)) < TF32
        assert rel_err(pg[id(sc.weight)], sd.weight.grad) < TF32
        assert rel_err(pg[id(sc.bias)], sd.bias.grad) < FP32 * 10
    if res is not None:
        assert rel_err(rv.g, _nhwc(rd.grad)) < FP32


def test_wgrad_split_matches_single_launch(cuda):
    from ivideogpt_b200 import ops
    from ivideogpt_b200.vq_model.train_plan import TokenizerTrainGraph
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(96, 16384, generator=g), torch.randn(200, 16384, generator=g)
    got = TokenizerTrainGraph._wgrad(a.to(cuda), b.to(cuda))
    assert rel_err(got, a.double() @ b.double().t()) < TF32
    assert rel_err(got, ops.gemm(a.to(cuda), b.to(cuda))) < 1e-4      # same TF32 products, different fp32 summation order


def test_linear_bwd_row_slices(cuda):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(5)
    M, K, Cc = 512, 128, 128
    W_ = nn.Parameter(torch.randn(3 * Cc, K, generator=g) * 0.1)
    b_ = nn.Parameter(torch.randn(3 * Cc, generator=g))
    x, res, dy = torch.randn(M, K, generator=g), torch.randn(M, Cc, generator=g), torch.randn(M, Cc, generator=g)
    Wd, bd = W_.detach().double().requires_grad_(True), b_.detach().double().requires_grad_(True)
    xd, rd = x.double().requires_grad_(True), res.double().requires_grad_(True)
    (F.linear(xd, Wd[Cc:2 * Cc], bd[Cc:2 * Cc]) + rd).backward(dy.double())
    gr = _graph()
    Wc, bc = nn.Parameter(W_.data.to(cuda)), nn.Parameter(b_.data.to(cuda))
    xv, rv = Var(x.to(cuda)), Var(res.to(cuda))
    out = gr.linear(xv, Wc, bc, Cc, 2 * Cc, residual=rv, tag="mha_k")
    pg = _run(gr, out, dy.to(cuda))
    assert rel_err(xv.g, xd.grad) < TF32 and rel_err(rv.g, rd.grad) < FP32
    assert rel_err(pg[id(Wc)], Wd.grad) < TF32 and rel_err(pg[id(bc)], bd.grad) < FP32 * 10
    assert not pg[id(Wc)][:Cc].any() and not pg[id(Wc)][2 * Cc:].any()


@pytest.mark.parametrize("heads,bdiv", [(1, 1), (4, 3)])
def test_attention_bwd(cuda, heads, bdiv):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(6)
    Fk, Lq, Lk, Cc = 2, 256, 512 if bdiv > 1 else 256, 128
    Fq = Fk * bdiv
    q, k, v = torch.randn(Fq, Lq, Cc, generator=g), torch.randn(Fk, Lk, Cc, generator=g), torch.randn(Fk, Lk, Cc, generator=g)
    do = torch.randn(Fq, Lq, Cc, generator=g)
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    dh = Cc // heads
    qh = qd.view(Fq, Lq, heads, dh).transpose(1, 2)
    kh = kd.repeat_interleave(bdiv, 0).view(Fq, Lk, heads, dh).transpose(1, 2)
    vh = vd.repeat_interleave(bdiv, 0).view(Fq, Lk, heads, dh).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1)
    o = (p @ vh).transpose(1, 2).reshape(Fq, Lq, Cc)
    o.backward(do.double())
    gr = _graph()
    qv, kv, vv = Var(q.to(cuda)), Var(k.to(cuda)), Var(v.to(cuda))
    out = gr.attn(qv, kv, vv, heads, bdiv)
    assert rel_err(out.v, o.detach()) < TF32
    _run(gr, out, do.to(cuda))
    assert rel_err(qv.g, qd.grad) < 2 * TF32 and rel_err(kv.g, kd.grad) < 2 * TF32 and rel_err(vv.g, vd.grad) < 2 * TF32


def test_upsample_patchify_bwd(cuda):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 32, 8, 8, generator=g)
    dy = torch.randn(3, 32, 16, 16, generator=g)
    xd = x.double().requires_grad_(True)
    F.interpolate(xd, scale_factor=2.0, mode="nearest").backward(dy.double())
    gr = _graph()
    xv = Var(_nhwc(x).to(cuda))
    out = gr.upsample(xv)
    _run(gr, out, _nhwc(dy).to(cuda))
    assert rel_err(xv.g, _nhwc(xd.grad)) < FP32
    # patchify / unpatchify are permutations: backward(forward-of-gradient) must be the inverse permutation
    gr = _graph()
    lat = torch.randn(2, 16, 16, 8, generator=g)
    lv = Var(lat.to(cuda))
    pv = gr.patchify(lv, 4)
    w = torch.randn(pv.v.shape, generator=g)
    _run(gr, pv, w.to(cuda))
    from ivideogpt_b200 import ops
    assert torch.equal(ops.patchify(lv.g, 4).cpu(), w)
    gr = _graph()
    uv_in = Var(w.to(cuda))
    uv = gr.unpatchify(uv_in, 4, 2, 16, 8)
    _run(gr, uv, lat.to(cuda))
    assert torch.equal(uv_in.g.cpu(), ops.patchify(lat.to(cuda), 4).cpu())


def test_vq_bwd_vs_oracle_quantizer(cuda):
    from oracle.vq_model_ref import RefVectorQuantizer
    from ivideogpt_b200.vq_model.modules import Codebook
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(8)
    N, D, K = 512, 64, 256
    ref = RefVectorQuantizer(K, D)
    ref.embedding.weight.data = torch.randn(K, D, generator=g)
    z = torch.randn(N, D, generator=g)
    dout = torch.randn(N, D, generator=g)
    zr = z.clone().requires_grad_(True)
    zq, loss, idx = ref(zr.t().reshape(1, D, N, 1))                       # NCHW with (H, W) = (N, 1): rows = z
    (zq.reshape(D, N).t() * dout).sum().backward(retain_graph=True)
    (0.7 * loss).backward()
    cb = Codebook(K, D)
    cb.embedding.weight.data = ref.embedding.weight.data.clone()
    cb = cb.to(cuda)
    gr = _graph()
    zv = Var(z.to(cuda))
    out, lossv = gr.vq(zv, cb, "ctx")
    assert torch.equal(gr.vq_indices["ctx"].cpu(), idx)
    assert abs(float(lossv.v) - float(loss)) < 1e-5 * float(loss)
    gr.backward_from([(out, dout.to(cuda)), (lossv, torch.tensor(0.7, device=cuda))])
    assert rel_err(zv.g, zr.grad) < 1e-5
    assert rel_err(gr.pgrads[id(cb.embedding.weight)][1], ref.embedding.weight.grad) < 1e-5


def test_conv_in_and_conv_out_bwd(cuda):
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(9)
    N, H, W, C0 = 2, 32, 32, 64
    px = torch.rand(N, 3, H, W, generator=g)
    cin = nn.Conv2d(3, C0, 3, padding=1)
    dy = torch.randn(N, C0, H, W, generator=g)
    cd = nn.Conv2d(3, C0, 3, padding=1).double()
    cd.load_state_dict({k: v.double() for k, v in cin.state_dict().items()})
    cd(px.double()).backward(dy.double())
    gr = _graph()
    cin = cin.to(cuda)
    out = gr.conv_in(px.to(cuda), cin)
    pg = _run(gr, out, _nhwc(dy).to(cuda))
    assert rel_err(pg[id(cin.weight)], cd.weight.grad) < TF32 and rel_err(pg[id(cin.bias)], cd.bias.grad) < FP32 * 10
    # GroupNorm + SiLU + conv C -> 3
    x = torch.randn(N, C0, H, W, generator=g)
    norm, cout = nn.GroupNorm(4, C0, eps=1e-6), nn.Conv2d(C0, 3, 3, padding=1)
    norm.weight.data, norm.bias.data = torch.randn(C0, generator=g), torch.randn(C0, generator=g) * 0.2
    dout = torch.randn(N, 3, H, W, generator=g)
    nd, od = nn.GroupNorm(4, C0, eps=1e-6).double(), nn.Conv2d(C0, 3, 3, padding=1).double()
    nd.load_state_dict({k: v.double() for k, v in norm.state_dict().items()})
    od.load_state_dict({k: v.double() for k, v in cout.state_dict().items()})
    xd = x.double().requires_grad_(True)
    y = od(F.silu(nd(xd)))
    y.backward(dout.double())
    gr = _graph()
    norm, cout = norm.to(cuda), cout.to(cuda)
    xv = Var(_nhwc(x).to(cuda))
    out = gr.conv_out(xv, norm, cout, torch.empty(N, 3, H, W, device=cuda))
    assert rel_err(out.v, y.detach()) < 1e-4          # the 3-channel conv is an fp32 FMA kernel
    pg = _run(gr, out, dout.to(cuda))
    assert rel_err(xv.g, _nhwc(xd.grad)) < TF32
    assert rel_err(pg[id(cout.weight)], od.weight.grad) < TF32 and rel_err(pg[id(cout.bias)], od.bias.grad) < FP32 * 10
    assert rel_err(pg[id(norm.weight)], nd.weight.grad) < TF32 and rel_err(pg[id(norm.bias)], nd.bias.grad) < TF32


def test_attention_dropout_same_mask(cuda):
    """Train mode: dropout on the attention probabilities (nn.MultiheadAttention(dropout=0.1), conditional_vae.py:24).  torch's
    RNG stream cannot be reproduced, so the float64 reference applies the PRODUCT's mask (regenerated from the seed with
    ops.dropout on a tensor of ones); forward and all three gradients must then agree, and the keep rate must be 1 - p."""
    from ivideogpt_b200 import ops
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(10)
    heads, bdiv, Fk, Lq, Lk, Cc, p_drop = 4, 2, 2, 256, 512, 128, 0.1
    Fq = Fk * bdiv
    q, k, v = torch.randn(Fq, Lq, Cc, generator=g), torch.randn(Fk, Lk, Cc, generator=g), torch.randn(Fk, Lk, Cc, generator=g)
    do = torch.randn(Fq, Lq, Cc, generator=g)
    gr = _graph()
    gr.training, gr.base_seed = True, 1234567
    qv, kv, vv = Var(q.to(cuda)), Var(k.to(cuda)), Var(v.to(cuda))
    out = gr.attn(qv, kv, vv, heads, bdiv, p_drop)
    seed = (gr.base_seed + 0x9E3779B97F4A7C15 * 1) & ((1 << 63) - 1)
    mask = ops.dropout(torch.ones(Fq * heads, Lq, Lk, device=cuda), p_drop, seed).double().cpu()      # 0 or 1/(1-p)
    assert abs(float((mask > 0).double().mean()) - (1 - p_drop)) < 2e-3
    assert set(mask.unique().tolist()) == {0.0, float(torch.tensor(1.0 / (1.0 - p_drop), dtype=torch.float32))}
    qd, kd, vd = (t.double().requires_grad_(True) for t in (q, k, v))
    dh = Cc // heads
    qh = qd.view(Fq, Lq, heads, dh).transpose(1, 2)
    kh = kd.repeat_interleave(bdiv, 0).view(Fq, Lk, heads, dh).transpose(1, 2)
    vh = vd.repeat_interleave(bdiv, 0).view(Fq, Lk, heads, dh).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1) * mask.view(Fq, heads, Lq, Lk)
    o = (p @ vh).transpose(1, 2).reshape(Fq, Lq, Cc)
    o.backward(do.double())
    assert rel_err(out.v, o.detach()) < TF32
    _run(gr, out, do.to(cuda))
    assert rel_err(qv.g, qd.grad) < 2 * TF32 and rel_err(kv.g, kd.grad) < 2 * TF32 and rel_err(vv.g, vd.grad) < 2 * TF32


def test_residual_dropout_site(cuda):
    """u = z + dropout(a): the gradient reaches `a` through the same mask and `z` untouched."""
    from ivideogpt_b200 import ops
    from ivideogpt_b200.vq_model.train_plan import Var
    g = torch.Generator().manual_seed(11)
    a, zt, du = torch.randn(300, 64, generator=g), torch.randn(300, 64, generator=g), torch.randn(300, 64, generator=g)
    gr = _graph()
    gr.training, gr.base_seed = True, 99
    av, zv = Var(a.to(cuda)), Var(zt.to(cuda))
    u = gr.add(zv, gr.dropout(av, 0.1))
    seed = (99 + 0x9E3779B97F4A7C15) & ((1 << 63) - 1)
    mask = ops.dropout(torch.ones(300, 64, device=cuda), 0.1, seed).cpu()
    assert rel_err(u.v, zt + a * mask) < FP32
    _run(gr, u, du.to(cuda))
    assert rel_err(av.g, du * mask) < FP32 and torch.equal(zv.g.cpu(), du)


@pytest.mark.parametrize("N,H,W,Ci,Co", [(4, 32, 32, 64, 128), (6, 16, 16, 128, 64), (2, 64, 64, 32, 32)])
def test_wgrad_conv3_without_im2col(cuda, N, H, W, Ci, Co):
    """Weight gradient from the padded channel-major copies (nine K-offset GEMMs) == the explicit-im2col route == float64."""
    from ivideogpt_b200 import ops
    from ivideogpt_b200.vq_model.train_plan import TokenizerTrainGraph
    g = torch.Generator().manual_seed(12)
    x, dy = torch.randn(N, H, W, Ci, generator=g), torch.randn(N, H, W, Co, generator=g)
    cols = F.unfold(x.permute(0, 3, 1, 2).double(), 3, padding=1)                                # [N, Ci*9, HW]
    want = torch.einsum("npo,nkp->ok", dy.double().view(N, H * W, Co), cols.view(N, Ci * 9, H * W))
    want = want.view(Co, Ci, 9).permute(0, 2, 1).reshape(Co, 9 * Ci)                              # (tap, ci) column order
    got = TokenizerTrainGraph._wgrad_conv3(dy.to(cuda), x.to(cuda))
    assert rel_err(got, want) < TF32
    old = TokenizerTrainGraph._wgrad(ops.transpose(dy.view(-1, Co).to(cuda)), ops.im2col3x3_t(x.to(cuda), 1))
    assert rel_err(got, old) < 1e-4


def test_transpose_pad_layout(cuda):
    from ivideogpt_b200 import ops
    g = torch.Generator().manual_seed(13)
    N, H, W, Cc = 3, 16, 16, 40
    x = torch.randn(N, H, W, Cc, generator=g)
    Wp = 20
    pad = F.pad(x.permute(3, 0, 1, 2), (1, Wp - W - 1, 1, 1))                                     # [C, N, H+2, Wp]
    out = ops.transpose_pad(x.to(cuda)).cpu()
    simg = out.shape[1] // N
    assert simg % 64 == 0 and simg >= (H + 2) * Wp + 4
    want = torch.zeros(Cc, N, simg)
    want[:, :, : (H + 2) * Wp] = pad.reshape(Cc, N, -1)
    assert torch.equal(out, want.view(Cc, -1))
    out3 = ops.transpose_pad(x.to(cuda), copies=3).cpu().view(3, Cc, N * simg)
    flat = want.view(Cc, -1)
    for b in range(3):                      # copy b read at q gives the unshifted array at q + (b - 1)
        assert torch.equal(out3[b], torch.roll(flat, -(b - 1), 1))
