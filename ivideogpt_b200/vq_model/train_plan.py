"""Training graph of the compressive tokenizer: forward with a tape, backward on the sm_100a kernels.

What the reference gets from torch autograd over `CompressiveVQModel.forward` (compressive_vq_model.py:332-369 + decode()
:290-330; `accelerator.backward(loss)` at train_tokenizer.py:734) is written out here: every forward op records a closure
that turns the gradient of its output into gradients of its inputs and parameters, and `backward()` runs the closures in
reverse.  The arithmetic is fp32 storage / TF32 tensor cores (the reference's own GPU arithmetic for this model):

  conv3x3 (+ fused 1x1 shortcut / residual)   dX = the SAME tcgen05 conv kernel on tap-flipped weights (stride 2: on the
                                              zero-inserted dY), dW = dY^T x im2col(X)^T on the tcgen05 GEMM, db = column sums
  GroupNorm (+ SiLU, + positional embedding)  csrc/tok_train.cu (two-stage deterministic reductions)
  Linear / 1x1 conv, attention products       tcgen05 GEMM on transposed operands; softmax backward from llama_train.cu
  cross-attention K / V shared by the frames  per-frame dK / dV summed over the frames of a clip (reduce_mid)
  VQ                                          straight-through dz + commitment terms, codebook rows via the embedding scatter

Nothing here falls back to torch arithmetic: torch allocates, views and zero-fills; gradient accumulation is add_to_f32.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import F32
from .plan import TokenizerPlan

_CTX_RES, _DYN_RES = 16, 4


class Var:
    """A forward value and the slot its gradient accumulates into."""
    __slots__ = ("v", "g", "needs")

    def __init__(self, v: torch.Tensor, needs: bool = True):
        self.v, self.g, self.needs = v, None, needs


class TokenizerTrainGraph:
    def __init__(self, model, plan: TokenizerPlan):
        if plan.dtype != torch.float32:
            raise NotImplementedError("tokenizer training runs in fp32 storage / TF32 tensor cores; "
                                      "call set_compute_dtype(torch.float32) (bf16 is an inference arithmetic here)")
        self.m, self.plan, self.pw = model, plan, plan.pw
        self.tape: List = []          # (closure, parameters whose gradient the closure produces)
        self.vars: List[Var] = []
        self.pgrads: Dict[int, Tuple[torch.nn.Parameter, torch.Tensor]] = {}
        self.vq_indices: Dict[str, torch.Tensor] = {}
        self.idx_override: Dict[str, torch.Tensor] = {}
        # 3x3 weight gradients from padded channel-major copies (see _wgrad_conv3); IVGPT_TOK_WGRAD_IM2COL=1 selects the
        # explicit transposed im2col instead (9x the activation; measured 35 % of the step's kernel time)
        self.wgrad_no_im2col = os.environ.get("IVGPT_TOK_WGRAD_IM2COL", "0") != "1"
        # train mode: the cross-attention dropouts of conditional_vae.py:24-25,52 are active (counter-based masks; the seed
        # comes from torch's CPU generator, so torch.manual_seed() fixes a run)
        self.training = bool(model.training) if model is not None else False
        self.base_seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if self.training else 0
        self._drop_sites = 0

    # ---- gradient slots ---------------------------------------------------------------------------------
    def _var(self, v: torch.Tensor, needs: bool = True) -> Var:
        var = Var(v, needs)
        self.vars.append(var)
        return var

    def _push(self, bwd, outs, *params):
        self.tape.append((bwd, tuple(p for p in params if p is not None), outs))

    def _acc(self, var: Optional[Var], g: torch.Tensor, own: bool):
        """own: g was produced for this slot alone (it may be kept and later added into in place)."""
        if var is None or not var.needs:
            return
        g = g.view(var.v.shape)
        if var.g is None:
            var.g = g if own else g.clone()
        else:
            ops.add_to_f32(var.g, g)

    def _pacc(self, p: torch.nn.Parameter, g: torch.Tensor, lo: Optional[int] = None):
        if not p.requires_grad:
            return
        key = id(p)
        if key not in self.pgrads:
            if lo is None and g.numel() == p.numel() and g.is_contiguous() and g._base is None:
                self.pgrads[key] = (p, g.view(p.shape))        # fresh tensor produced for this parameter: keep it
                return
            self.pgrads[key] = (p, torch.zeros(p.shape, dtype=torch.float32, device=g.device))
        buf = self.pgrads[key][1]
        tgt = buf if lo is None else buf[lo: lo + g.shape[0]]
        g = g.contiguous()
        assert tgt.numel() == g.numel(), (tuple(tgt.shape), tuple(g.shape))
        ops.add_to_f32(tgt, g)

    # ---- ops ----------------------------------------------------------------------------------------------
    def view(self, x: Var, shape) -> Var:
        out = self._var(x.v.view(shape), x.needs)

        def bwd():
            if out.g is not None:
                self._acc(x, out.g, own=True)
        self._push(bwd, (out,))
        return out

    @staticmethod
    def _wgrad(dyT: torch.Tensor, xT: torch.Tensor) -> torch.Tensor:
        """[Cout, K] = dY^T [Cout, P] x X^T [K, P]^T.  The contraction runs over the P pixels: with only a handful of
        output tiles, the pixels are cut into `ks` slices (a batch dimension whose stride is a K offset), one GEMM launch
        fills [ks, Cout, K] partials and reduce_mid adds them in a fixed order."""
        M, P = dyT.shape
        N = xT.shape[0]
        tiles = ((M + 127) // 128) * ((N + 127) // 128)
        ks = 1
        while ks * 2 * tiles <= 296 and P % (ks * 2 * 32) == 0 and P // (ks * 2) >= 1024:
            ks *= 2
        if ks == 1:
            return ops.gemm(dyT, xT)
        Kc = P // ks
        part = torch.empty(ks, M, N, dtype=torch.float32, device=dyT.device)
        ops.gemm_raw(ops.gemm_desc(
            dtype=F32, a=dyT.data_ptr(), lda=dyT.stride(0), a_bstride=Kc, a_rows=M, a_cols=Kc, a_batches=ks,
            b=xT.data_ptr(), ldb=xT.stride(0), b_bstride=Kc, b_rows=N, b_cols=Kc, b_batches=ks,
            M=M, N=N, K=Kc, batch=ks, heads=1, a_bsel=2, b_bsel=2, o_bsel=2,
            out=part.data_ptr(), ldo=N, out_bstride=M * N, out_dtype=F32))
        return ops.reduce_mid(part, 1, ks).view(M, N)

    @staticmethod
    def _wgrad_conv3(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """Weight gradient [Cout, 9*Cin] of a 3x3 / stride-1 / pad-1 conv from dY [N,H,W,Cout] and X [N,H,W,Cin] WITHOUT an
        im2col copy: both tensors are transposed into channel-major arrays with a zero frame around every image
        (ops.transpose_pad, row pitch Wp); tap (a, b) is then the GEMM dY^T x X^T over the pixel axis with X^T read at the
        constant K offset (a-1)*Wp + (b-1), and the frame of zeros in dY^T makes the image borders come out right.  The row
        part of the offset is a TMA box coordinate (out-of-range columns read as zeros); the column part cannot be -- a box
        must start on a 16-byte boundary of the innermost dimension -- so X^T is written three times, shifted by b-1.
        Three launches (one per kernel row a; the column taps ride the descriptor's head index: B rows h*Cin of the stacked
        copies, output columns h*Cin), the pixel axis cut into whole-image slices as the batch so that every launch has
        enough tiles, partials added in fixed order."""
        N, H, W, Co = dy.shape
        Ci = x.shape[-1]
        dyT, xT3 = ops.transpose_pad(dy), ops.transpose_pad(x, copies=3)
        Pp = dyT.shape[1]
        simg = Pp // N
        Wp = (W + 2 + 3) // 4 * 4
        tiles = ((Co + 127) // 128) * ((Ci + 127) // 128) * 3
        ks = 1
        for d in range(1, N + 1):                       # largest divisor of N that keeps the launch within ~2 waves
            if N % d == 0 and d * tiles <= 296:
                ks = d
        Kc = (N // ks) * simg
        part = torch.empty(ks, Co, 9 * Ci, dtype=torch.float32, device=dy.device)
        for a in range(3):
            ops.gemm_raw(ops.gemm_desc(
                dtype=F32, a=dyT.data_ptr(), lda=Pp, a_bstride=Kc, a_rows=Co, a_cols=Kc, a_batches=ks,
                b=xT3.data_ptr(), ldb=Pp, b_bstride=Kc, b_rows=3 * Ci, b_cols=Kc, b_batches=ks,
                M=Co, N=Ci, K=Kc, batch=ks * 3, heads=3, a_bsel=1, a_bdiv=1, b_bsel=1, b_bdiv=1, o_bsel=1,
                b_kbase=(a - 1) * Wp, b_nhead=Ci, o_nhead=Ci,
                out=part.data_ptr() + a * 3 * Ci * 4, ldo=9 * Ci, out_bstride=Co * 9 * Ci, out_dtype=F32))
        return part[0] if ks == 1 else ops.reduce_mid(part, 1, ks).view(Co, 9 * Ci)

    def conv(self, x: Var, conv, stride: int = 1, shortcut=None, x2: Optional[Var] = None,
             residual: Optional[Var] = None) -> Var:
        w, b = self.pw.conv3(conv, torch.float32, shortcut=shortcut)
        y = ops.conv3x3(x.v, w, b, stride=stride, x2=None if x2 is None else x2.v,
                        residual=None if residual is None else residual.v)
        out = self._var(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            N, Ho, Wo, Co = dy.shape
            P = N * Ho * Wo
            Ci = x.v.shape[-1]
            dy2 = dy.view(P, Co)
            if conv.bias.requires_grad or (shortcut is not None and shortcut.bias.requires_grad):
                db = ops.colsum(dy2)
                self._pacc(conv.bias, db)
                if shortcut is not None:
                    self._pacc(shortcut.bias, db.clone())        # two parameters must not share one gradient tensor
            need_sw = shortcut is not None and shortcut.weight.requires_grad
            fast = stride == 1 and self.wgrad_no_im2col
            dyT = ops.transpose(dy2) if ((conv.weight.requires_grad and not fast) or need_sw) else None      # [Co, P]
            if conv.weight.requires_grad:
                dWp = self._wgrad_conv3(dy, x.v) if fast else self._wgrad(dyT, ops.im2col3x3_t(x.v, stride))   # [Co, 9*Ci]
                self._pacc(conv.weight, dWp.view(Co, 3, 3, Ci).permute(0, 3, 1, 2))
            if shortcut is not None:
                C2 = x2.v.shape[-1]
                if need_sw:
                    self._pacc(shortcut.weight, self._wgrad(dyT, ops.transpose(x2.v.view(P, C2))).view(Co, C2, 1, 1))
                if x2.needs:
                    self._acc(x2, ops.gemm(dy2, self.pw.linear_t(shortcut.weight)), own=True)
            if residual is not None:
                self._acc(residual, dy, own=False)
            if x.needs:
                src = ops.zero_insert2x(dy) if stride == 2 else dy
                self._acc(x, ops.conv3x3(src, self.pw.conv3_dgrad(conv), None), own=True)
        self._push(bwd, (out,), conv.weight, conv.bias, *((shortcut.weight, shortcut.bias) if shortcut is not None else ()))
        return out

    def gn(self, x: Var, norm, silu: bool, samples: Optional[int] = None, pos=None) -> Var:
        n = x.v.shape[0] if samples is None else samples
        gamma, beta = self.pw.f32(norm.weight), self.pw.f32(norm.bias)
        stats = ops.groupnorm_stats(x.v, n, norm.num_groups, norm.eps)
        y = ops.groupnorm_apply(x.v, stats, gamma, beta, silu, None if pos is None else self.pw.f32(pos))
        out = self._var(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            dx, dg, db = ops.groupnorm_bwd(x.v, dy, stats, gamma, beta, silu, want_dx=x.needs)
            self._pacc(norm.weight, dg)
            self._pacc(norm.bias, db)
            if pos is not None and pos.requires_grad:
                self._pacc(pos, ops.reduce_mid(dy, 1, dy.numel() // pos.numel()).view(pos.shape))
            if x.needs:
                self._acc(x, dx, own=True)
        self._push(bwd, (out,), norm.weight, norm.bias, pos)
        return out

    def linear(self, x: Var, weight, bias, lo: Optional[int] = None, hi: Optional[int] = None,
               residual: Optional[Var] = None, tag: str = "lin") -> Var:
        """x [M, K] -> [M, N] with rows lo:hi of `weight` (all rows when lo is None); 1x1-conv weights are flattened."""
        if lo is None:
            w, b = self.pw.linear(weight, bias, torch.float32)
        else:
            w, b = self.pw.rows(weight, bias, lo, hi, torch.float32, tag)
        y = ops.gemm(x.v, w, b, residual=None if residual is None else residual.v)
        out = self._var(y)

        def bwd():
            dy = out.g
            if dy is None:
                return
            if bias is not None and bias.requires_grad:
                self._pacc(bias, ops.colsum(dy), lo)
            if weight.requires_grad:
                dW = self._wgrad(ops.transpose(dy), ops.transpose(x.v))                         # [N, K]
                self._pacc(weight, dW.view((dW.shape[0],) + tuple(weight.shape[1:])), lo)
            if residual is not None:
                self._acc(residual, dy, own=False)
            if x.needs:
                self._acc(x, ops.gemm(dy, self.pw.linear_t(weight, lo, hi)), own=True)
        self._push(bwd, (out,), weight, bias)
        return out

    def _next_seed(self) -> int:
        self._drop_sites += 1
        return (self.base_seed + 0x9E3779B97F4A7C15 * self._drop_sites) & ((1 << 63) - 1)

    def dropout(self, x: Var, p: float) -> Var:
        if p <= 0.0:
            return x
        seed = self._next_seed()
        out = self._var(ops.dropout(x.v, p, seed))

        def bwd():
            if out.g is not None:
                self._acc(x, ops.dropout(out.g, p, seed), own=True)
        self._push(bwd, (out,))
        return out

    def add(self, a: Var, b: Var) -> Var:
        out = self._var(ops.axpby(a.v, b.v, 1.0, 1.0))

        def bwd():
            if out.g is not None:
                self._acc(a, out.g, own=False)
                self._acc(b, out.g, own=True)
        self._push(bwd, (out,))
        return out

    def silu(self, x: Var) -> Var:
        out = self._var(ops.silu(x.v))

        def bwd():
            if out.g is not None:
                self._acc(x, ops.silu(x.v, out.g), own=True)
        self._push(bwd, (out,))
        return out

    # attention products: the two descriptor forms of plan.TokenizerPlan.attention
    @staticmethod
    def _scores(a, b, heads: int, bdiv: int, alpha: float) -> torch.Tensor:
        """a [Fa, M, C], b [Fb, N, C] -> [Fa*heads, M, N] fp32: per head h, a[f][:, h] @ b[f // bdiv][:, h]^T * alpha."""
        Fa, M, Cc = a.shape
        Fb, N, _ = b.shape
        dh = Cc // heads
        out = torch.empty(Fa * heads, M, N, dtype=torch.float32, device=a.device)
        ops.gemm_raw(ops.gemm_desc(
            dtype=F32, a=a.data_ptr(), lda=Cc, a_bstride=M * Cc, a_rows=M, a_cols=Cc, a_batches=Fa,
            b=b.data_ptr(), ldb=Cc, b_bstride=N * Cc, b_rows=N, b_cols=Cc, b_batches=Fb,
            M=M, N=N, K=dh, batch=Fa * heads, heads=heads, a_bsel=1, a_bdiv=1, b_bsel=1, b_bdiv=bdiv,
            o_bsel=2, a_khead=dh, b_khead=dh,
            out=out.data_ptr(), ldo=N, out_bstride=M * N, out_dtype=F32, alpha=alpha))
        return out

    @staticmethod
    def _values(p, bt, heads: int, bdiv: int) -> torch.Tensor:
        """p [Fa*heads, M, K], bt [Fb, C, K] -> [Fa, M, C]: head h of frame f gets p[f*heads+h] @ bt[f // bdiv][h*dh:(h+1)*dh]^T."""
        FH, M, K = p.shape
        Fb, Cc, _ = bt.shape
        Fa = FH // heads
        dh = Cc // heads
        assert p.is_contiguous() and bt.stride(2) == 1
        out = torch.empty(Fa, M, Cc, dtype=torch.float32, device=p.device)
        ops.gemm_raw(ops.gemm_desc(
            dtype=F32, a=p.data_ptr(), lda=K, a_bstride=M * K, a_rows=M, a_cols=K, a_batches=FH,
            b=bt.data_ptr(), ldb=bt.stride(1), b_bstride=bt.stride(0), b_rows=Cc, b_cols=K, b_batches=Fb,
            M=M, N=dh, K=K, batch=FH, heads=heads, a_bsel=2, b_bsel=1, b_bdiv=bdiv,
            o_bsel=1, b_nhead=dh, o_nhead=dh,
            out=out.data_ptr(), ldo=Cc, out_bstride=M * Cc, out_dtype=F32))
        return out

    def attn(self, q: Var, k: Var, v: Var, heads: int, bdiv: int, p_drop: float = 0.0) -> Var:
        """q [Fq, Lq, C]; k, v [Fk, Lk, C] with Fq = Fk * bdiv (frame f attends to clip f // bdiv) -> [Fq, Lq, C].
        p_drop: dropout on the attention probabilities (nn.MultiheadAttention(dropout=...), conditional_vae.py:24): P' =
        keep(seed, i) ? P / (1 - p) : 0 feeds P'.V; the backward regenerates the mask from the seed."""
        Fq, Lq, Cc = q.v.shape
        Fk, Lk, _ = k.v.shape
        scale = 1.0 / math.sqrt(Cc // heads)
        rows = Fq * heads * Lq
        s = self._scores(q.v, k.v, heads, bdiv, scale)
        p = torch.empty_like(s)
        ops.softmax(s, p, rows, Lq, Lk, Lk, Lk, False)
        del s
        seed = self._next_seed() if p_drop > 0.0 else 0
        pd = ops.dropout(p, p_drop, seed) if p_drop > 0.0 else p
        out = self._var(self._values(pd, ops.transpose(v.v), heads, bdiv))

        def bwd():
            do = out.g
            if do is None:
                return
            dp = self._scores(do, v.v, heads, bdiv, 1.0)                                  # [Fq*h, Lq, Lk]
            if p_drop > 0.0:
                ops.dropout(dp, p_drop, seed, out=dp)                                     # same mask, same 1/(1-p)
            ds = torch.empty_like(dp)
            ops.softmax_bwd(p, dp, ds, rows, Lq, Lk, Lk, False, scale)
            del dp
            if v.needs:
                dvf = self._values(ops.transpose(pd), ops.transpose(do), heads, 1)         # [Fq, Lk, C] per query frame
                self._acc(v, dvf if bdiv == 1 else ops.reduce_mid(dvf, Fk, bdiv), own=True)
            if q.needs:
                self._acc(q, self._values(ds, ops.transpose(k.v), heads, bdiv), own=True)
            if k.needs:
                dkf = self._values(ops.transpose(ds), ops.transpose(q.v), heads, 1)
                self._acc(k, dkf if bdiv == 1 else ops.reduce_mid(dkf, Fk, bdiv), own=True)
        self._push(bwd, (out,))
        return out

    def upsample(self, x: Var) -> Var:
        out = self._var(ops.upsample2x(x.v))

        def bwd():
            if out.g is not None:
                self._acc(x, ops.upsample2x_bwd(out.g), own=True)
        self._push(bwd, (out,))
        return out

    def patchify(self, x: Var, p: int) -> Var:
        F_, R, _, Cc = x.v.shape
        out = self._var(ops.patchify(x.v, p))

        def bwd():
            if out.g is not None:
                self._acc(x, ops.patchify(out.g, p, inverse=True, frames=F_, res=R, ch=Cc), own=True)
        self._push(bwd, (out,))
        return out

    def unpatchify(self, x: Var, p: int, frames: int, res: int, ch: int) -> Var:
        out = self._var(ops.patchify(x.v, p, inverse=True, frames=frames, res=res, ch=ch))

        def bwd():
            if out.g is not None:
                self._acc(x, ops.patchify(out.g, p), own=True)
        self._push(bwd, (out,))
        return out

    def vq(self, z: Var, codebook, name: str, beta: float = 1.0) -> Tuple[Var, Var]:
        """diffusers VectorQuantizer(legacy=False).forward: (straight-through z_q, beta*mean((sg zq - z)^2) + mean((zq - sg z)^2))."""
        cb = codebook.embedding.weight.detach().float().contiguous()
        idx = self.idx_override.get(name)
        if idx is None:
            idx = ops.vq_argmin(z.v, cb)
        self.vq_indices[name] = idx
        zq, loss = ops.vq_commit(z.v, cb, idx, torch.float32, beta=beta)
        out, lossv = self._var(zq), self._var(loss)

        def bwd():
            if out.g is None and lossv.g is None:
                return
            dz, de = ops.vq_bwd(z.v, zq, out.g, lossv.g, beta)
            if lossv.g is not None and codebook.embedding.weight.requires_grad:
                dE = torch.zeros_like(cb)
                ops.embed_bwd(idx, de, dE)
                self._pacc(codebook.embedding.weight, dE)
            self._acc(z, dz, own=True)
        self._push(bwd, (out, lossv), codebook.embedding.weight)
        return out, lossv

    def conv_in(self, px: torch.Tensor, conv) -> Var:
        """px [N,3,H,W] fp32 (no gradient) -> [N,H,W,C0]."""
        N, _, H, W = px.shape
        w27, b27 = self.pw.conv_in27(conv)
        out = self._var(ops.conv_in(px.view(1, N, 3, H, W), w27, b27, torch.float32, 0, N))

        def bwd():
            dy = out.g
            if dy is None:
                return
            Co = dy.shape[-1]
            dy2 = dy.view(-1, Co)
            if conv.bias.requires_grad:
                self._pacc(conv.bias, ops.colsum(dy2))
            if conv.weight.requires_grad:
                colT = ops.im2col3x3_t(ops.nchw_to_nhwc(px), 1, k_rows=32)             # rows (tap, ci), padded 27 -> 32
                dW = self._wgrad(ops.transpose(dy2), colT)[:, :27]
                self._pacc(conv.weight, dW.reshape(Co, 3, 3, 3).permute(0, 3, 1, 2))
        self._push(bwd, (out,), conv.weight, conv.bias)
        return out

    def conv_out(self, x: Var, norm, conv, out_nchw: torch.Tensor) -> Var:
        """GroupNorm + SiLU + the C -> 3 conv, written as NCHW frames (the fused forward kernel of the inference path)."""
        N, H, W, Cc = x.v.shape
        gamma, beta = self.pw.f32(norm.weight), self.pw.f32(norm.bias)
        stats = ops.groupnorm_stats(x.v, N, norm.num_groups, norm.eps)
        w3, b3 = self.pw.conv_out3(conv)
        ops.conv_out3(x.v, stats, gamma, beta, w3, b3, out_nchw.view(1, N, 3, H, W), 0, N)
        out = self._var(out_nchw)

        def bwd():
            d = out.g
            if d is None:
                return
            dy32 = ops.nchw_to_nhwc(d.contiguous(), 32)                                  # 3 real channels + 29 zero ones
            dy2 = dy32.view(-1, 32)
            if conv.bias.requires_grad:
                self._pacc(conv.bias, ops.colsum(dy2)[:3])
            if conv.weight.requires_grad:
                y = ops.groupnorm_apply(x.v, stats, gamma, beta, True)                   # recomputed, not kept from the forward
                dWp = self._wgrad_conv3(dy32, y) if self.wgrad_no_im2col else self._wgrad(ops.transpose(dy2), ops.im2col3x3_t(y, 1))   # [32, 9*C]
                del y
                self._pacc(conv.weight, dWp[:3].reshape(3, 3, 3, Cc).permute(0, 3, 1, 2))
            dyn = ops.conv3x3(dy32, self.pw.conv3_dgrad(conv, cout_pad=32), None)        # gradient of the normalised input
            dx, dg, db = ops.groupnorm_bwd(x.v, dyn, stats, gamma, beta, True, want_dx=x.needs)
            self._pacc(norm.weight, dg)
            self._pacc(norm.bias, db)
            if x.needs:
                self._acc(x, dx, own=True)
        self._push(bwd, (out,), conv.weight, conv.bias, norm.weight, norm.bias)
        return out

    # ---- blocks (mirror plan.TokenizerPlan) ------------------------------------------------------------
    def resnet(self, x: Var, r) -> Var:
        h = self.conv(self.gn(x, r.norm1, True), r.conv1)
        g = self.gn(h, r.norm2, True)
        if r.conv_shortcut is not None:
            return self.conv(g, r.conv2, shortcut=r.conv_shortcut, x2=x)
        return self.conv(g, r.conv2, residual=x)

    def cross_attention(self, z: Var, ctx_feat: Var, blk, clips: int) -> Var:
        F_, H, W, Cc = z.v.shape
        t = ctx_feat.v.shape[0] // clips
        assert blk.kv_frames == t, f"cross-attention built for {blk.kv_frames} context frames, got {t}"
        L, Lkv = H * W, t * H * W
        kv = self.view(self.gn(ctx_feat, blk.kv_norm, False, samples=clips, pos=blk.kv_pos_emb), (clips * Lkv, Cc))
        q = self.view(self.gn(z, blk.q_norm, False, pos=blk.q_pos_emb), (F_ * L, Cc))
        W_, b_ = blk.att.in_proj_weight, blk.att.in_proj_bias
        qp = self.view(self.linear(q, W_, b_, 0, Cc, tag="mha_q"), (F_, L, Cc))
        kp = self.view(self.linear(kv, W_, b_, Cc, 2 * Cc, tag="mha_k"), (clips, Lkv, Cc))
        vp = self.view(self.linear(kv, W_, b_, 2 * Cc, 3 * Cc, tag="mha_v"), (clips, Lkv, Cc))
        p_drop = float(getattr(blk, "dropout", 0.0)) if self.training else 0.0
        o = self.view(self.attn(qp, kp, vp, blk.heads, F_ // clips, p_drop), (F_ * L, Cc))
        z2 = self.view(z, (F_ * L, Cc))
        if p_drop > 0.0:                                                                 # resid_dropout, conditional_vae.py:52
            u = self.add(z2, self.dropout(self.linear(o, blk.att.out_proj.weight, blk.att.out_proj.bias), p_drop))
        else:
            u = self.linear(o, blk.att.out_proj.weight, blk.att.out_proj.bias, residual=z2)
        return self.view(self.silu(u), (F_, H, W, Cc))                                   # silu(z + attn), conditional_vae.py:55

    def mid_attention(self, x: Var, a) -> Var:
        F_, H, W, Cc = x.v.shape
        L = H * W
        tok = self.view(self.gn(x, a.group_norm, False), (F_ * L, Cc))
        qp = self.view(self.linear(tok, a.to_q.weight, a.to_q.bias), (F_, L, Cc))
        kp = self.view(self.linear(tok, a.to_k.weight, a.to_k.bias), (F_, L, Cc))
        vp = self.view(self.linear(tok, a.to_v.weight, a.to_v.bias), (F_, L, Cc))
        o = self.view(self.attn(qp, kp, vp, 1, 1), (F_ * L, Cc))
        out = self.linear(o, a.to_out[0].weight, a.to_out[0].bias, residual=self.view(x, (F_ * L, Cc)))
        return self.view(out, (F_, H, W, Cc))

    def mid(self, x: Var, m) -> Var:
        x = self.resnet(x, m.resnets[0])
        if m.has_attention:
            x = self.mid_attention(x, m.attentions[0])
        return self.resnet(x, m.resnets[1])

    def encode(self, px: torch.Tensor, enc, ctx_feats: Optional[List[Var]] = None, clips: int = 0):
        x = self.conv_in(px, enc.conv_in)
        feats = [x]
        k = 0
        for i, stage in enumerate(enc.down_blocks):
            for r in stage.resnets:
                x = self.resnet(x, r)
            if stage.downsamplers is not None:
                x = self.conv(x, stage.downsamplers[0].conv, stride=2)
            if ctx_feats is not None and x.v.shape[2] <= enc.max_att_resolution:
                x = self.cross_attention(x, ctx_feats[i + 1], enc.cross_att_blocks[k], clips)
                k += 1
            feats.append(x)
        x = self.mid(x, enc.mid_block)
        feats.append(x)
        return self.conv(self.gn(x, enc.conv_norm_out, True), enc.conv_out), feats

    def decode(self, latent: Var, dec, out_nchw: Optional[torch.Tensor], ctx_feats: Optional[List[Var]] = None, clips: int = 0):
        x = self.conv(latent, dec.conv_in)
        feats = [x]
        x = self.mid(x, dec.mid_block)
        feats.append(x)
        if ctx_feats is not None:
            x = self.cross_attention(x, ctx_feats[1], dec.cross_att_blocks[0], clips)
        for i, stage in enumerate(dec.up_blocks):
            for r in stage.resnets:
                x = self.resnet(x, r)
            if stage.upsamplers is not None:
                x = self.conv(self.upsample(x), stage.upsamplers[0].conv)
            if ctx_feats is not None and x.v.shape[2] <= dec.max_att_resolution:
                x = self.cross_attention(x, ctx_feats[i + 2], dec.cross_att_blocks[i + 1], clips)
            feats.append(x)
        if out_nchw is None:        # the caller runs GroupNorm + SiLU + conv_out as a separate autograd node (see forward_train)
            return x, feats
        return self.conv_out(x, dec.conv_norm_out, dec.conv_out, out_nchw), feats

    # ---- the whole graph -----------------------------------------------------------------------------------
    def forward(self, sample: torch.Tensor, dyn_sample: torch.Tensor, segment_len: int):
        """compressive_vq_model.py:332-369 / :290-330.  Returns Vars (x_last, ref_dec, commit_loss, dyn_commit_loss) where x_last
        is the input of the conditional decoder's last layer (conv_norm_out -> SiLU -> conv_out): that layer is a separate
        autograd node (head()), so that a gradient asked w.r.t. it alone does not sweep the whole tape."""
        m = self.m
        t, f, cr = m.context_length, int(segment_len), _CTX_RES
        B = sample.shape[0] // t
        Cc, H, W = sample.shape[1:]
        dev = sample.device
        ctx = sample.detach().to(torch.float32).contiguous()
        fut = dyn_sample.detach().to(torch.float32).contiguous()
        h, feats = self.encode(ctx, m.encoder)
        z_ctx = self.linear(self.view(h, (-1, h.v.shape[-1])), m.quant_conv.weight, m.quant_conv.bias)
        zq_c, commit = self.vq(z_ctx, m.quantize, "ctx")
        lat_c = self.view(self.linear(zq_c, m.post_quant_conv.weight, m.post_quant_conv.bias), (B * t, cr, cr, m.latent_channels))
        ref_dec, dec_feats = self.decode(lat_c, m.decoder, torch.empty(B * t, m.config["out_channels"], H, W, dtype=torch.float32, device=dev))
        d, _ = self.encode(fut, m.cond_encoder, ctx_feats=feats, clips=B)
        z_dyn = self.linear(self.patchify(d, m.patch_size), m.quant_linear.weight, m.quant_linear.bias)
        zq_d, dyn_commit = self.vq(z_dyn, m.dynamics_quantize, "dyn")
        pd = self.linear(zq_d, m.post_quant_linear.weight, m.post_quant_linear.bias)
        lat_d = self.unpatchify(pd, m.patch_size, B * f, cr, m.latent_channels)
        x_last, _ = self.decode(lat_d, m.cond_decoder, None, ctx_feats=dec_feats, clips=B)
        self.outputs = (x_last, ref_dec, commit, dyn_commit)
        return self.outputs

    def head(self, x_last: torch.Tensor) -> Var:
        """The conditional decoder's last layer on a detached activation: (input Var, output Var)."""
        m = self.m
        N, H, W, _ = x_last.shape
        self.head_in = self._var(x_last)
        out = self.conv_out(self.head_in, m.cond_decoder.conv_norm_out, m.cond_decoder.conv_out,
                            torch.empty(N, m.config["out_channels"], H, W, dtype=torch.float32, device=x_last.device))
        self.outputs = (out,)
        return out

    # The closures capture the graph, and the graph owning its tape would be a reference cycle: the activations of a finished
    # step would then live until the cyclic garbage collector runs (measured: 56 GiB held and a 4x slower step at 4 clips).
    # The autograd node owns the tape instead (ctx.tape) and lends it back for the duration of a backward sweep, so everything
    # a step saved is released by reference counting the moment autograd drops the node.
    def detach_tape(self):
        t = (self.tape, self.vars)
        self.tape, self.vars = [], []
        return t

    def attached(self, t):
        import contextlib

        @contextlib.contextmanager
        def lend():
            self.tape, self.vars = t
            try:
                yield
            finally:
                self.tape, self.vars = [], []
        return lend()

    def backward_from(self, seeds, needed=None, want=()) -> Dict[int, Tuple[torch.nn.Parameter, torch.Tensor]]:
        """seeds: [(Var, gradient tensor)].  Runs the tape in reverse and returns {id(param): (param, grad fp32)}.  The tape is
        kept: the graph can be differentiated again (train_tokenizer.py:706-707 takes torch.autograd.grad(..., retain_graph=
        True) of two losses w.r.t. the last decoder layer before the real backward; that layer is its own autograd node, see
        forward_train).  `needed` (ids of the parameters whose gradient is wanted) stops the sweep at the earliest op that
        touches one of them; `want`: Vars whose gradient is kept in self.wanted."""
        for var in self.vars:
            var.g = None
        self.pgrads = {}
        for var, g in seeds:
            if g is not None:
                var.g = g.detach().to(torch.float32).contiguous().view(var.v.shape).clone()
        lo = 0
        if needed is not None:
            hit = [i for i, (_, ps, _) in enumerate(self.tape) if any(id(p) in needed for p in ps)]
            lo = min(hit) if hit else len(self.tape)
        for fn, _, outs in reversed(self.tape[lo:]):
            fn()
            for var in outs:              # this op's output gradient has been consumed: release it now, not at the end
                var.g = None
        self.wanted = [var.g for var in want]
        for var in self.vars:
            var.g = None
        return self.pgrads

    def backward(self, d_dec, d_ref_dec, d_commit, d_dyn_commit, needed=None):
        return self.backward_from(list(zip(self.outputs, (d_dec, d_ref_dec, d_commit, d_dyn_commit))), needed)


class _TokenizerTrainFn(torch.autograd.Function):
    """autograd boundary of the body: (x_last, ref_dec, commit_loss, dyn_commit_loss) = f(parameters); the inputs are pixels."""

    @staticmethod
    def forward(ctx, graph: TokenizerTrainGraph, sample, dyn_sample, segment_len, *params):
        with torch.no_grad():
            x_last, ref_dec, commit, dyn_commit = graph.forward(sample, dyn_sample, segment_len)
        ctx.graph, ctx.params = graph, params
        ctx.tape = graph.detach_tape()
        ctx.set_materialize_grads(False)
        # fresh tensor objects: autograd hangs grad_fn (-> this node -> the tape) on what is returned, and the graph keeps its
        # own Vars -- returning those very objects would tie graph -> tensor -> node -> graph into a cycle that no collector
        # sees through (measured: one step's activations leaked per step)
        return x_last.v.detach(), ref_dec.v.detach(), commit.v.detach(), dyn_commit.v.detach()

    @staticmethod
    def backward(ctx, d_x_last, d_ref_dec, d_commit, d_dyn_commit):
        graph = ctx.graph
        with torch.no_grad(), torch.cuda.device(graph.outputs[0].v.device), graph.attached(ctx.tape):
            pg = graph.backward(d_x_last, d_ref_dec, d_commit, d_dyn_commit)
        grads = []
        for p, need in zip(ctx.params, ctx.needs_input_grad[4:]):
            e = pg.get(id(p)) if need else None
            grads.append(None if e is None else e[1].to(p.dtype))
        graph.pgrads = {}
        return (None, None, None, None, *grads)


class _TokenizerHeadFn(torch.autograd.Function):
    """dec = conv_out(silu(conv_norm_out(x_last))) of the conditional decoder as its own node: torch.autograd.grad(loss,
    cond_decoder.conv_out.weight, retain_graph=True) (train_tokenizer.py:706-707, the adaptive GAN weight) runs this node only."""

    @staticmethod
    def forward(ctx, graph: TokenizerTrainGraph, x_last, *params):
        with torch.no_grad():
            out = graph.head(x_last.detach())
        ctx.graph, ctx.params = graph, params
        ctx.tape = graph.detach_tape()
        ctx.set_materialize_grads(False)
        return out.v.detach()

    @staticmethod
    def backward(ctx, d_dec):
        graph = ctx.graph
        if d_dec is None:
            return (None, None) + (None,) * len(ctx.params)
        with torch.no_grad(), torch.cuda.device(graph.outputs[0].v.device), graph.attached(ctx.tape):
            pg = graph.backward_from([(graph.outputs[0], d_dec)], want=(graph.head_in,))
        grads = []
        for p, need in zip(ctx.params, ctx.needs_input_grad[2:]):
            e = pg.get(id(p)) if need else None
            grads.append(None if e is None else e[1].to(p.dtype))
        graph.pgrads = {}
        dx, graph.wanted = (graph.wanted[0] if ctx.needs_input_grad[1] else None), []
        return (None, dx, *grads)


def forward_train(model, plan: TokenizerPlan, sample, dyn_sample, segment_len, idx_override=None):
    """(dec, ref_dec, commit_loss, dyn_commit_loss) with autograd history, and the body graph (for its VQ indices)."""
    graph = TokenizerTrainGraph(model, plan)
    if idx_override:
        graph.idx_override = dict(idx_override)
    cd = model.cond_decoder
    head_params = [cd.conv_norm_out.weight, cd.conv_norm_out.bias, cd.conv_out.weight, cd.conv_out.bias]
    head_ids = {id(p) for p in head_params}
    params = [p for p in model.parameters() if p.requires_grad and id(p) not in head_ids]
    x_last, ref_dec, commit, dyn_commit = _TokenizerTrainFn.apply(graph, sample, dyn_sample, segment_len, *params)
    head = TokenizerTrainGraph(model, plan)
    dec = _TokenizerHeadFn.apply(head, x_last, *head_params)
    return (dec, ref_dec, commit, dyn_commit), graph
