"""Thin torch-tensor front end over the C ABI (ivideogpt_b200/_lib.py).

torch is used here only as the owner of device memory and streams: every function checks that its tensors
live on a CUDA device, allocates the output with torch.empty and enqueues one or more of OUR kernels on
torch.cuda.current_stream().  Nothing in this file computes with torch ops.
"""
from __future__ import annotations

import ctypes as C
import functools
import types
from typing import Optional

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_SILU, ACT_SWIGLU, BF16, F32, ConvDesc, GemmDesc


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"ivideogpt_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}")


def _cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.B200LibraryError("ivideogpt_b200 runs on CUDA (sm_100a) only; got a tensor on "
                                        f"{t.device}.  There is no CPU fallback.")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def torch_dtype(code: int):
    return torch.bfloat16 if code == BF16 else torch.float32


# ------------------------------------------------------------------------------------------------
# VQ argmin
# ------------------------------------------------------------------------------------------------
def vq_argmin(z: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """idx[n] = argmin_k ||z_n - e_k||  (int64).  z [N,64] fp32, codebook [K,64] fp32."""
    _cuda(z, codebook)
    assert z.dtype == torch.float32 and codebook.dtype == torch.float32
    z = z.contiguous()
    codebook = codebook.contiguous()
    N, D = z.shape
    K = codebook.shape[0]
    idx = torch.empty(N, dtype=torch.int64, device=z.device)
    enorm = torch.empty(K, dtype=torch.float32, device=z.device)
    packed = torch.empty(max(N, 1), dtype=torch.int64, device=z.device)
    lib = _lib.load()
    _lib.check(lib.ivgpt_vq_argmin(z.data_ptr(), codebook.data_ptr(), enorm.data_ptr(), packed.data_ptr(),
                                   idx.data_ptr(), N, K, D, _stream()), "vq_argmin")
    return idx


# ------------------------------------------------------------------------------------------------
# GEMM family
# ------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
         residual: Optional[torch.Tensor] = None, act: int = ACT_NONE, out_dtype: Optional[torch.dtype] = None,
         out: Optional[torch.Tensor] = None, alpha: float = 1.0, bn: int = 0) -> torch.Tensor:
    """out[M,N] = act(alpha * a[M,K] @ w[N,K]^T + bias + residual).  a, w share dtype (fp32->TF32, or bf16)."""
    _cuda(a, w, bias, residual, out)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1], (a.shape, w.shape)
    assert a.dtype == w.dtype
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        out = torch.empty(M, n_out, dtype=out_dtype or a.dtype, device=a.device)
    assert out.stride(1) == 1 and out.shape[0] == M and out.shape[1] == n_out
    d = GemmDesc()
    d.dtype = _dt(a); d.bn = bn
    d.a = a.data_ptr(); d.lda = a.stride(0); d.a_rows = M; d.a_cols = K; d.a_batches = 1
    d.b = w.data_ptr(); d.ldb = w.stride(0); d.b_rows = N; d.b_cols = K; d.b_batches = 1
    d.M, d.N, d.K = M, N, K
    d.batch = 1; d.heads = 1; d.a_bdiv = 1; d.b_bdiv = 1
    d.out = out.data_ptr(); d.ldo = out.stride(0); d.out_dtype = _dt(out)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        d.bias = bias.data_ptr()
    if residual is not None:
        assert residual.shape == out.shape and residual.stride(1) == 1
        d.residual = residual.data_ptr(); d.ldr = residual.stride(0); d.res_dtype = _dt(residual)
    d.act = act; d.alpha = alpha
    _lib.check(_lib.load().ivgpt_gemm(C.byref(d), _stream()), "gemm")
    return out


def gemm_desc(**kw) -> GemmDesc:
    d = GemmDesc()
    d.batch = 1; d.heads = 1; d.a_bdiv = 1; d.b_bdiv = 1; d.alpha = 1.0
    for k, v in kw.items():
        setattr(d, k, v)
    return d


def gemm_raw(d: GemmDesc):
    _lib.check(_lib.load().ivgpt_gemm(C.byref(d), _stream()), "gemm")


def conv3x3(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], stride: int = 1,
            x2: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, act: int = ACT_NONE,
            out_dtype: Optional[torch.dtype] = None, bn: int = 0, gn_groups: int = 0, gn_in=None) -> torch.Tensor:
    """3x3 conv over NHWC x [N,H,W,Cin] with packed weights [Cout, 9*Cin (+C2)]; optional fused 1x1 source x2.
    gn_in = (scale [N,Cin], shift [N,Cin], silu): the GroupNorm (+ SiLU) that precedes the conv, applied to the operand tiles
    inside the kernel (groupnorm_coeff() makes the coefficients) -- the normalised activation never exists in HBM.
    gn_groups > 0: the epilogue also emits GroupNorm partial statistics of the output; they ride on the returned tensor
    as `out.gn_part = (partials [N, slabs, G, 2], slabs)` for groupnorm_stats_from_parts()."""
    _cuda(x, w_packed, bias, x2, residual)
    assert x.is_contiguous() and w_packed.is_contiguous() and x.dtype == w_packed.dtype
    N, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    C2 = 0 if x2 is None else x2.shape[-1]
    assert w_packed.shape[1] == 9 * Cin + C2, (w_packed.shape, Cin, C2)
    Ho, Wo = H // stride, W // stride
    out = torch.empty(N, Ho, Wo, Cout, dtype=out_dtype or x.dtype, device=x.device)
    d = ConvDesc()
    d.dtype = _dt(x); d.bn = bn
    d.x = x.data_ptr(); d.N, d.Hin, d.Win, d.Cin, d.stride = N, H, W, Cin, stride
    d.w = w_packed.data_ptr(); d.Cout = Cout
    if x2 is not None:
        assert x2.is_contiguous() and x2.shape[:3] == (N, Ho, Wo) and x2.dtype == x.dtype
        d.x2 = x2.data_ptr(); d.C2 = C2
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout
        d.bias = bias.data_ptr()
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
        d.residual = residual.data_ptr(); d.res_dtype = _dt(residual)
    d.act = act
    d.out = out.data_ptr(); d.out_dtype = _dt(out)
    if gn_in is not None:
        sc, sh, silu = gn_in
        assert stride == 1 and sc.dtype == torch.float32 and sh.dtype == torch.float32 and sc.shape == (N, Cin) == sh.shape
        assert sc.is_contiguous() and sh.is_contiguous()
        d.in_scale = sc.data_ptr(); d.in_shift = sh.data_ptr(); d.in_silu = int(bool(silu))
    part = None
    if gn_groups > 0:
        bn_c, slabs_c = C.c_int(0), C.c_int(0)
        _lib.check(_lib.load().ivgpt_conv3x3_plan(C.byref(d), C.byref(bn_c), C.byref(slabs_c)), "conv3x3_plan")
        d.bn = bn_c.value
        part = torch.empty(N, slabs_c.value, gn_groups, 2, dtype=torch.float32, device=x.device)
        d.gn_part = part.data_ptr(); d.gn_groups = gn_groups
    _lib.check(_lib.load().ivgpt_conv3x3(C.byref(d), _stream()), "conv3x3")
    if part is not None:
        out.gn_part = (part, slabs_c.value)
    return out


# ------------------------------------------------------------------------------------------------
# GroupNorm and friends
# ------------------------------------------------------------------------------------------------
def groupnorm_stats(x: torch.Tensor, samples: int, groups: int, eps: float) -> torch.Tensor:
    """x is any contiguous [..., C] tensor viewed as [samples, rows, C]; returns (mean, rstd) [samples, G, 2]."""
    _cuda(x)
    assert x.is_contiguous()
    Cc = x.shape[-1]
    rows = x.numel() // (samples * Cc)
    slabs = (rows + 63) // 64
    part = torch.empty(samples * slabs * groups * 2, dtype=torch.float32, device=x.device)
    stats = torch.empty(samples, groups, 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_groupnorm_stats(_dt(x), x.data_ptr(), part.data_ptr(), stats.data_ptr(), samples,
                                                 rows, Cc, groups, eps, _stream()), "groupnorm_stats")
    return stats


def groupnorm_stats_from_parts(x: torch.Tensor, samples: int, groups: int, eps: float) -> torch.Tensor:
    """Statistics of a conv output whose epilogue already produced partial sums (x.gn_part); `samples` may merge
    several consecutive frames (joint kv_norm)."""
    part, slabs = x.gn_part
    frames = part.shape[0]
    assert frames % samples == 0 and part.shape[2] == groups
    per = frames // samples
    Cc = x.shape[-1]
    count = float(x.numel() // (samples * Cc)) * (Cc // groups)
    stats = torch.empty(samples, groups, 2, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_groupnorm_finalize(part.data_ptr(), stats.data_ptr(), samples, slabs * per, groups,
                                                    count, eps, _stream()), "groupnorm_finalize")
    return stats


def groupnorm_coeff(stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor):
    """(scale, shift) [samples, C] fp32 with y = x * scale + shift == GroupNorm(x) for statistics [samples, G, 2]."""
    _cuda(stats, gamma, beta)
    samples, groups = stats.shape[0], stats.shape[1]
    Cc = gamma.numel()
    scale = torch.empty(samples, Cc, dtype=torch.float32, device=stats.device)
    shift = torch.empty(samples, Cc, dtype=torch.float32, device=stats.device)
    _lib.check(_lib.load().ivgpt_groupnorm_coeff(stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), scale.data_ptr(),
                                                 shift.data_ptr(), samples, Cc, groups, _stream()), "groupnorm_coeff")
    return scale, shift


def groupnorm_apply(x: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, silu: bool,
                    pos: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _cuda(x, stats, gamma, beta, pos, out)
    assert x.is_contiguous()
    Cc = x.shape[-1]
    samples, groups = stats.shape[0], stats.shape[1]
    total_rows = x.numel() // Cc
    y = torch.empty_like(x) if out is None else out
    assert y.is_contiguous() and y.shape == x.shape and y.dtype == x.dtype
    pos_rows = 0
    if pos is not None:
        assert pos.dtype == torch.float32 and pos.is_contiguous() and pos.shape[1] == Cc
        pos_rows = pos.shape[0]
    coef = torch.empty(2 * samples * Cc, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_groupnorm_apply(_dt(x), x.data_ptr(), y.data_ptr(), stats.data_ptr(),
                                                 gamma.data_ptr(), beta.data_ptr(), _ptr(pos), coef.data_ptr(), total_rows,
                                                 total_rows // samples, Cc, groups, int(silu), pos_rows, _stream()),
               "groupnorm_apply")
    return y


def conv_in(clips: torch.Tensor, w27: torch.Tensor, bias: torch.Tensor, dtype: torch.dtype, frame_offset: int,
            frames_per_clip: int) -> torch.Tensor:
    """clips [B,T,3,H,W] fp32 contiguous; processes frames [frame_offset, frame_offset+frames_per_clip) of each clip."""
    _cuda(clips, w27, bias)
    assert clips.dtype == torch.float32 and clips.is_contiguous() and clips.dim() == 5 and clips.shape[2] == 3
    B, T, _, H, W = clips.shape
    N = B * frames_per_clip
    Cout = w27.shape[0]
    y = torch.empty(N, H, W, Cout, dtype=dtype, device=clips.device)
    _lib.check(_lib.load().ivgpt_conv_in(_dt(y), clips.data_ptr(), w27.data_ptr(), bias.data_ptr(), y.data_ptr(), N,
                                         H, W, Cout, frames_per_clip, T, frame_offset, _stream()), "conv_in")
    return y


def conv_out3(x: torch.Tensor, stats: torch.Tensor, gamma, beta, w_packed: torch.Tensor, bias: torch.Tensor,
              out_clips: torch.Tensor, frame_offset: int, frames_per_clip: int):
    """Writes frames into out_clips [B,T,3,H,W] fp32 at slots [frame_offset, frame_offset+frames_per_clip)."""
    _cuda(x, stats, gamma, beta, w_packed, bias, out_clips)
    N, H, W, Cc = x.shape
    assert out_clips.is_contiguous() and out_clips.dtype == torch.float32
    _lib.check(_lib.load().ivgpt_conv_out3(_dt(x), x.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                           w_packed.data_ptr(), bias.data_ptr(), out_clips.data_ptr(), N, H, W, Cc,
                                           stats.shape[1], frames_per_clip, out_clips.shape[1], frame_offset,
                                           _stream()), "conv_out3")
    return out_clips


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    _cuda(x)
    N, H, W, Cc = x.shape
    y = torch.empty(N, 2 * H, 2 * W, Cc, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().ivgpt_upsample2x(_dt(x), x.data_ptr(), y.data_ptr(), N, H, W, Cc, _stream()), "upsample2x")
    return y


def patchify(x: torch.Tensor, p: int, inverse: bool = False, frames: int = 0, res: int = 0, ch: int = 0):
    _cuda(x)
    assert x.is_contiguous()
    if not inverse:
        F_, R, _, Cc = x.shape
        y = torch.empty(F_ * (R // p) ** 2, p * p * Cc, dtype=x.dtype, device=x.device)
    else:
        F_, R, Cc = frames, res, ch
        y = torch.empty(F_, R, R, Cc, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().ivgpt_patchify(_dt(x), x.data_ptr(), y.data_ptr(), F_, R, Cc, p, int(inverse), _stream()),
               "patchify")
    return y


def convert(x: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    _cuda(x)
    if x.dtype == dtype:
        return x
    assert x.is_contiguous()
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    _lib.check(_lib.load().ivgpt_convert(_dt(x), x.data_ptr(), _dt(y), y.data_ptr(), x.numel(), _stream()), "convert")
    return y


def vq_commit(z: torch.Tensor, codebook: torch.Tensor, idx: torch.Tensor, dtype: torch.dtype, beta: float = 1.0):
    """(z_q [N, D] in `dtype`, loss 0-d fp32) = rows of the codebook picked by idx and (beta + 1) * mean((z_q - z)^2)."""
    _cuda(z, codebook, idx)
    assert z.dtype == torch.float32 and codebook.dtype == torch.float32 and idx.dtype == torch.int64
    z, codebook, idx = z.contiguous(), codebook.contiguous(), idx.contiguous()
    N, D = z.shape
    zq = torch.empty(N, D, dtype=dtype, device=z.device)
    part = torch.empty(296, dtype=torch.float32, device=z.device)
    loss = torch.empty((), dtype=torch.float32, device=z.device)
    _lib.check(_lib.load().ivgpt_vq_commit(_dt(zq), z.data_ptr(), codebook.data_ptr(), idx.data_ptr(), zq.data_ptr(), N, D,
                                           codebook.shape[0], float(beta), part.data_ptr(), loss.data_ptr(), _stream()),
               "vq_commit")
    return zq, loss


def tokens_serialise(idx_ctx, idx_dyn, B, t, f, cr, dr, n_vq, n_dyn, want_labels=True):
    _cuda(idx_ctx, idx_dyn)
    L = t * (cr + 1) - 1 + f * (dr + 1)
    tokens = torch.empty(B, L, dtype=torch.int64, device=idx_ctx.device)
    labels = torch.empty(B, L, dtype=torch.int64, device=idx_ctx.device) if want_labels else None
    _lib.check(_lib.load().ivgpt_tokens_serialise(idx_ctx.data_ptr(), idx_dyn.data_ptr(), tokens.data_ptr(),
                                                  _ptr(labels), B, t, f, cr, dr, n_vq, n_dyn, _stream()),
               "tokens_serialise")
    return tokens, labels


def tokens_gather(tokens, cb_ctx, cb_dyn, t, f, cr, dr, dtype, bad_ctx=None):
    """bad_ctx: optional int32 [1] device flag, set to 1 by the kernel when a context position holds an id outside the
    context codebook (the reference's embedding lookup raises there, compressive_vq_model.py:238)."""
    _cuda(tokens, cb_ctx, cb_dyn, bad_ctx)
    assert tokens.dtype == torch.int64 and tokens.is_contiguous()
    B, L = tokens.shape
    D = cb_ctx.shape[1]
    qc = torch.empty(B * t * cr, D, dtype=dtype, device=tokens.device)
    qd = torch.empty(B * f * dr, D, dtype=dtype, device=tokens.device)
    _lib.check(_lib.load().ivgpt_tokens_gather(_dt(qc), tokens.data_ptr(), cb_ctx.data_ptr(), cb_dyn.data_ptr(),
                                               qc.data_ptr(), qd.data_ptr(), B, t, f, cr, dr, D, cb_ctx.shape[0],
                                               cb_dyn.shape[0], L, _ptr(bad_ctx), _stream()), "tokens_gather")
    return qc, qd


# ------------------------------------------------------------------------------------------------
# Llama pieces (raw-pointer style: the engine in transformer/engine.py owns the buffers)
# ------------------------------------------------------------------------------------------------
def embed(ids, ids_stride, L, dpos, table, x, M):
    _lib.check(_lib.load().ivgpt_embed(ids.data_ptr(), ids_stride, L, _ptr(dpos), table.data_ptr(), x.data_ptr(), M,
                                       table.shape[1], table.shape[0], _stream()), "embed")


def add_rows(x, e):
    _lib.check(_lib.load().ivgpt_add_rows(x.data_ptr(), e.data_ptr(), x.numel(), _stream()), "add_rows")


def rmsnorm(x, w, y, M, eps):
    _lib.check(_lib.load().ivgpt_rmsnorm(_dt(y), x.data_ptr(), w.data_ptr(), y.data_ptr(), M, x.shape[-1], eps,
                                         _stream()), "rmsnorm")


def rope_kv(qkv, q_out, k_cache, v_cache_t, B, Lq, heads, Lmax, pos0, dpos, cos_tab, sin_tab, v_rows=None):
    _lib.check(_lib.load().ivgpt_rope_kv(_dt(qkv), qkv.data_ptr(), q_out.data_ptr(), k_cache.data_ptr(),
                                         v_cache_t.data_ptr(), _ptr(v_rows), B, Lq, heads, Lmax, pos0, _ptr(dpos),
                                         cos_tab.data_ptr(), sin_tab.data_ptr(), _stream()), "rope_kv")


def softmax(S, P, rows, Lq, Lk, lds, ldp, causal, causal_off=0):
    _lib.check(_lib.load().ivgpt_softmax(_dt(P), S.data_ptr(), P.data_ptr(), rows, Lq, Lk, lds, ldp, int(causal),
                                         causal_off, _stream()), "softmax")


def flash_attn(q, k, vt, out, B, heads, Lq, Lk, k_bstride, vt_bstride, vt_ld, causal=True, scale=0.125, lse=None):
    """Fused causal attention (bf16, head_dim 64): q [B,heads,Lq,64], k rows [.., Lk, 64] with batch pitch k_bstride, vt = V^T
    rows [.., 64, vt_ld] with batch pitch vt_bstride; out [B*Lq, heads*64].  Scores / probabilities never reach HBM."""
    _cuda(q, k, vt, out, lse)
    assert q.dtype == torch.bfloat16 and k.dtype == torch.bfloat16 and vt.dtype == torch.bfloat16 and out.dtype == torch.bfloat16
    assert q.is_contiguous() and out.stride(-1) == 1
    _lib.check(_lib.load().ivgpt_flash_attn(q.data_ptr(), k.data_ptr(), vt.data_ptr(), out.data_ptr(), _ptr(lse), B, heads, Lq, Lk,
                                            Lq * 64, k_bstride, vt_bstride, vt_ld, out.stride(0), int(causal), float(scale),
                                            _stream()), "flash_attn")
    return out


def decode_attn(q, k_cache, v_cache_t, out, B, heads, Lmax, Lcur, dpos, scale):
    _lib.check(_lib.load().ivgpt_decode_attn(_dt(q), q.data_ptr(), k_cache.data_ptr(), v_cache_t.data_ptr(),
                                             out.data_ptr(), B, heads, Lmax, Lcur, _ptr(dpos), scale, _stream()),
               "decode_attn")


def decode_attn_fused(qkv, k_cache, v_cache_t, out, B, heads, Lmax, pos, dpos, cos_tab, sin_tab, scale):
    _lib.check(_lib.load().ivgpt_decode_attn_fused(_dt(qkv), qkv.data_ptr(), k_cache.data_ptr(), v_cache_t.data_ptr(),
                                                   out.data_ptr(), B, heads, Lmax, pos, _ptr(dpos), cos_tab.data_ptr(),
                                                   sin_tab.data_ptr(), scale, _stream()), "decode_attn_fused")


def set_deterministic(on: bool):
    """Force the single tcgen05.mma issuing warp (the default) even when two issuers were opted in."""
    _lib.load().ivgpt_set_deterministic(int(bool(on)))


def set_gemm_mh2(on: bool):
    """Opt in to 256 x 256 CTA tiles for BN = 256 launches (measured slower than 128 x 256; kept tested)."""
    _lib.load().ivgpt_set_gemm_mh2(int(bool(on)))


def set_mma_issuers(n: int):
    """1 (default) or 2 tcgen05.mma issuing warps in the GEMM / conv kernel (2: summation order of the last bits not fixed)."""
    _lib.load().ivgpt_set_mma_issuers(int(n))


def set_pdl(on: bool):
    _lib.load().ivgpt_set_pdl(int(on))


def argmax(logits, ld, rows, V, out, out_stride, dpos=None, out_offset=0):
    _lib.check(_lib.load().ivgpt_argmax(logits.data_ptr(), ld, rows, V, out.data_ptr() + 8 * out_offset, out_stride,
                                        _ptr(dpos), _stream()), "argmax")


def topk_sample(logits, ld, rows, V, k, temperature, seed, step, out, out_stride, dpos=None, out_offset=0, dseed=None):
    _lib.check(_lib.load().ivgpt_topk_sample(logits.data_ptr(), ld, rows, V, k, temperature, seed, step,
                                             out.data_ptr() + 8 * out_offset, out_stride, _ptr(dpos), _ptr(dseed),
                                             _stream()), "topk_sample")


def ce_loss(logits, ld, B, L, V, labels):
    """Shifted cross-entropy over logits [B,L,ld] fp32 / labels [B,L] int64; returns (mean loss 0-d, per-row losses,
    count [1] = number of labelled positions, on the device)."""
    _cuda(logits, labels)
    assert labels.dtype == torch.int64 and labels.is_contiguous()
    rows = B * (L - 1)
    loss_rows = torch.empty(max(rows, 1), dtype=torch.float32, device=logits.device)
    valid = torch.empty(max(rows, 1), dtype=torch.float32, device=logits.device)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    _lib.check(_lib.load().ivgpt_ce_loss(logits.data_ptr(), ld, B, L, V, labels.data_ptr(), loss_rows.data_ptr(),
                                         valid.data_ptr(), out.data_ptr(), _stream()), "ce_loss")
    return out[0], loss_rows[:rows], out[1:2]


def incr(p, by=1):
    _lib.check(_lib.load().ivgpt_incr(p.data_ptr(), by, _stream()), "incr")


def slot_embed_add(x, slot_emb, dpos, B, hidden, slot0, period):
    """x[b] += slot_emb[b, i] when the position being fed (*dpos) is forced slot i (action_model.py:80-81)."""
    assert slot_emb.dtype == torch.float32 and slot_emb.is_contiguous() and slot_emb.shape[0] == B
    _lib.check(_lib.load().ivgpt_slot_embed_add(x.data_ptr(), slot_emb.data_ptr(), dpos.data_ptr(), B, hidden, slot0, period,
                                                slot_emb.shape[1], _stream()), "slot_embed_add")


def slot_force(tokens, dpos, B, slot0, period, token):
    """tokens[b, *dpos + 1] = token when that position is a forced slot (action_model.py:109-110)."""
    _lib.check(_lib.load().ivgpt_slot_force(tokens.data_ptr(), tokens.stride(0), dpos.data_ptr(), B, slot0, period,
                                            int(token), _stream()), "slot_force")


# ------------------------------------------------------------------------------------------------
# training (backward) pieces
# ------------------------------------------------------------------------------------------------
def transpose(x: torch.Tensor, out: Optional[torch.Tensor] = None, pad_to: int = 8) -> torch.Tensor:
    """Batched 2-D transpose of the last two dims: x [..., R, C] (last dim contiguous) -> [..., C, Rp] with
    Rp = R rounded up to `pad_to` (TMA row pitches must be multiples of 16 bytes); returns the [..., C, :R] view."""
    _cuda(x, out)
    assert x.stride(-1) == 1
    R, Cc = x.shape[-2], x.shape[-1]
    batch = x.numel() // (R * Cc) if x.dim() > 2 else 1
    if x.dim() > 2:
        lead = x.shape[:-2]
        assert x.is_contiguous() or x.dim() == 3, "batched transpose needs a contiguous (or 3-D strided) input"
        bs_in = x.stride(-3) if x.dim() >= 3 else R * x.stride(-2)
        if x.dim() > 3:
            assert x.is_contiguous()
    else:
        lead, bs_in = (), 0
    Rp = (R + pad_to - 1) // pad_to * pad_to
    if out is None:
        # the pad columns [R, Rp) are read by the consumer GEMM as part of its K range and must be zero; with no pad columns
        # the zero fill would be a wasted pass (5 % of the train64 step's kernel time in profiles/r02/launches_train64.txt)
        out = (torch.empty if Rp == R else torch.zeros)(*lead, Cc, Rp, dtype=x.dtype, device=x.device)
    _lib.check(_lib.load().ivgpt_transpose(_dt(x), x.data_ptr(), out.data_ptr(), batch, R, Cc, x.stride(-2), Rp, bs_in,
                                           Cc * Rp, _stream()), "transpose")
    return out[..., :R]


def swiglu(gu: torch.Tensor, dact: Optional[torch.Tensor] = None) -> torch.Tensor:
    """forward: act [M,I] from interleaved gu [M,2I]; backward (dact given): d_gu [M,2I]."""
    _cuda(gu, dact)
    assert gu.is_contiguous()
    M, I2 = gu.shape
    if dact is None:
        out = torch.empty(M, I2 // 2, dtype=gu.dtype, device=gu.device)
        _lib.check(_lib.load().ivgpt_swiglu(_dt(gu), 0, gu.data_ptr(), None, out.data_ptr(), M * (I2 // 2), _stream()), "swiglu")
    else:
        assert dact.is_contiguous() and dact.dtype == gu.dtype
        out = torch.empty_like(gu)
        _lib.check(_lib.load().ivgpt_swiglu(_dt(gu), 1, gu.data_ptr(), dact.data_ptr(), out.data_ptr(), M * (I2 // 2), _stream()), "swiglu_bwd")
    return out


def rmsnorm_bwd(x: torch.Tensor, w: torch.Tensor, dy: torch.Tensor, dres: torch.Tensor, eps: float) -> torch.Tensor:
    """dres (fp32, in place) += d/dx of rmsnorm; returns dw (fp32 [H])."""
    _cuda(x, w, dy, dres)
    M, H = x.shape
    assert x.dtype == torch.float32 and dres.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous()
    part = torch.empty((M + 7) // 8, H, dtype=torch.float32, device=x.device)
    dw = torch.empty(H, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_rmsnorm_bwd(_dt(dy), x.data_ptr(), w.data_ptr(), dy.data_ptr(), dres.data_ptr(),
                                             part.data_ptr(), dw.data_ptr(), M, H, eps, _stream()), "rmsnorm_bwd")
    return dw


def softmax_bwd(P, dP, dS, rows, Lq, Lk, ld, causal, scale):
    _lib.check(_lib.load().ivgpt_softmax_bwd(_dt(P), P.data_ptr(), dP.data_ptr(), dS.data_ptr(), rows, Lq, Lk, ld,
                                             int(causal), scale, _stream()), "softmax_bwd")


def rope_bwd(dq, dk, dv, dqkv, B, L, heads, cos_tab, sin_tab):
    _lib.check(_lib.load().ivgpt_rope_bwd(_dt(dqkv), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dqkv.data_ptr(), B, L,
                                          heads, cos_tab.data_ptr(), sin_tab.data_ptr(), _stream()), "rope_bwd")


def ce_bwd(logits, ld, B, L, V, labels, count, gscale, dlogits):
    _lib.check(_lib.load().ivgpt_ce_bwd(_dt(dlogits), logits.data_ptr(), ld, B, L, V, labels.data_ptr(), count.data_ptr(),
                                        gscale, dlogits.data_ptr(), dlogits.stride(-2), _stream()), "ce_bwd")


def embed_bwd(ids, dx, dE):
    _lib.check(_lib.load().ivgpt_embed_bwd(ids.data_ptr(), dx.data_ptr(), dE.data_ptr(), ids.numel(), dE.shape[1],
                                           dE.shape[0], _stream()), "embed_bwd")


def adamw(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, gscale=1.0):
    """Fused AdamW update of `p` in place (raw-pointer kernel).  Pass the Parameter itself (or a tensor sharing its
    version counter, e.g. p.detach()), NOT p.data: the write is announced with increment_version so that every cache
    keyed on (data_ptr, _version) -- the packed kernel-layout weight copies of LlamaWeights / PackedWeights -- is
    rebuilt before the next forward."""
    _cuda(p, g, m, v)
    assert all(t.dtype == torch.float32 and t.is_contiguous() for t in (p, g, m, v))
    _lib.check(_lib.load().ivgpt_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                                       eps, weight_decay, step, gscale, _stream()), "adamw")
    torch.autograd.graph.increment_version(p)


def dropout(x: torch.Tensor, p: float, seed: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = keep(seed, i) ? x / (1 - p) : 0 over the flattened contiguous x (out may be x itself)."""
    _cuda(x, out)
    assert x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    assert out.is_contiguous() and out.dtype == x.dtype and out.numel() == x.numel()
    _lib.check(_lib.load().ivgpt_dropout(_dt(x), x.data_ptr(), out.data_ptr(), x.numel(), float(p),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, _stream()), "dropout")
    return out


def add_to_f32(y, x):
    _lib.check(_lib.load().ivgpt_add_to_f32(_dt(x), y.data_ptr(), x.data_ptr(), x.numel(), _stream()), "add_to_f32")


def transpose_raw(src: torch.Tensor, src_off: int, dst: torch.Tensor, batch: int, R: int, Cc: int, ld_in: int,
                  ld_out: int, bs_in: int, bs_out: int):
    """dst[b][c][r] = src[src_off + b*bs_in + r*ld_in + c]  (element offsets); dst row pitch ld_out, batch pitch bs_out."""
    es = src.element_size()
    _lib.check(_lib.load().ivgpt_transpose(_dt(src), src.data_ptr() + src_off * es, dst.data_ptr(), batch, R, Cc, ld_in,
                                           ld_out, bs_in, bs_out, _stream()), "transpose")


def preprocess_resize(frames: torch.Tensor, size_hw, channels_last: bool = True, divisor: float = 255.0) -> torch.Tensor:
    """`images / 255` + antialiased bilinear resize (inference/utils.py:12-16).  frames: CUDA uint8 or fp32, [T,H,W,C] when
    channels_last else [T,C,H,W] (any strides); returns fp32 [T, C, out_h, out_w]."""
    if not frames.is_cuda:
        raise RuntimeError("preprocess_resize requires a CUDA tensor (no CPU fallback)")
    if frames.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"preprocess_resize: uint8 or float32 frames expected, got {frames.dtype}")
    if frames.dim() != 4:
        raise ValueError("preprocess_resize: frames must be 4-D")
    if channels_last:
        T, H, W, Cc = frames.shape
        st, sy, sx, sc = frames.stride()
    else:
        T, Cc, H, W = frames.shape
        st, sc, sy, sx = frames.stride()
    oh, ow = int(size_hw[0]), int(size_hw[1])
    out = torch.empty(T, Cc, oh, ow, dtype=torch.float32, device=frames.device)
    _lib.check(_lib.load().ivgpt_preprocess_resize(2 if frames.dtype == torch.uint8 else 0, frames.data_ptr(), st, sy, sx, sc,
                                                   T, H, W, Cc, out.data_ptr(), oh, ow, float(divisor), _stream()),
               "preprocess_resize")
    return out


# ------------------------------------------------------------------------------------------------
# tokenizer training (backward) pieces -- fp32 NHWC (csrc/tok_train.cu)
# ------------------------------------------------------------------------------------------------
def _f32c(*ts):
    for t in ts:
        if t is not None:
            assert t.dtype == torch.float32 and t.is_contiguous(), (t.dtype, t.shape, t.stride())


def colsum(x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """out[c] (+)= sum over rows of x [M, C] (fp32, last dim contiguous): bias gradients."""
    _cuda(x, out)
    assert x.dim() == 2 and x.dtype == torch.float32 and x.stride(1) == 1
    M, Cc = x.shape
    if out is None:
        assert not accumulate
        out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    nb = 256          # row slabs of the first stage; the second stage adds them serially per column (592 was 37 us per call)
    part = torch.empty(nb * Cc, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_colsum(x.data_ptr(), M, Cc, x.stride(0), part.data_ptr(), nb, out.data_ptr(),
                                        int(accumulate), _stream()), "colsum")
    return out


def groupnorm_bwd(x: torch.Tensor, dy: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  silu: bool, want_dx: bool = True):
    """Backward of y = act(GroupNorm(x)) for x viewed [samples, rows, C] with statistics [samples, G, 2]:
    returns (dx or None, dgamma [C], dbeta [C])."""
    _cuda(x, dy, stats, gamma, beta)
    _f32c(x, dy, stats, gamma, beta)
    assert dy.shape == x.shape
    Cc = x.shape[-1]
    samples, groups = stats.shape[0], stats.shape[1]
    rows = x.numel() // (samples * Cc)
    lib = _lib.load()
    chunks = lib.ivgpt_groupnorm_bwd_chunks(samples, rows)
    ws = torch.empty(samples * (chunks + 1) * Cc * 2 + samples * groups * 2, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x) if want_dx else None
    dgamma = torch.empty(Cc, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(Cc, dtype=torch.float32, device=x.device)
    _lib.check(lib.ivgpt_groupnorm_bwd(x.data_ptr(), dy.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                       int(silu), samples, rows, Cc, groups, ws.data_ptr(), _ptr(dx), 0, dgamma.data_ptr(),
                                       dbeta.data_ptr(), 0, _stream()), "groupnorm_bwd")
    return dx, dgamma, dbeta


def im2col3x3_t(x: torch.Tensor, stride: int = 1, k_rows: Optional[int] = None) -> torch.Tensor:
    """colT [k_rows, N*Ho*Wo] fp32 with colT[tap*C + c][pixel] = the input pixel tap (a, b) of that output pixel sees
    (zero padding as the forward conv; rows beyond 9*C are zero): the K-major operand of the conv weight gradient."""
    _cuda(x)
    _f32c(x)
    N, H, W, Cc = x.shape
    k_rows = 9 * Cc if k_rows is None else k_rows
    P = N * (H // stride) * (W // stride)
    assert P % 4 == 0
    colT = torch.empty(k_rows, P, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_im2col3x3_t(x.data_ptr(), colT.data_ptr(), N, H, W, Cc, stride, k_rows, _stream()),
               "im2col3x3_t")
    return colT


def transpose_pad(x: torch.Tensor, copies: int = 1) -> torch.Tensor:
    """[N,H,W,C] fp32 -> [copies*C, N*img_stride]: channel-major with a zero frame around every image, row pitch Wp = W+2
    rounded up to 4, img_stride = (H+2)*Wp + 4 rounded up to 64.  copies=1: out[c, n*img_stride + (y+1)*Wp + (x+1)];
    copies=3: rows [b*C, (b+1)*C) hold the same data shifted by b-1 columns (out index minus (b-1))."""
    _cuda(x)
    _f32c(x)
    assert copies in (1, 3)
    N, H, W, Cc = x.shape
    Wp = (W + 2 + 3) // 4 * 4
    img_stride = ((H + 2) * Wp + 4 + 63) // 64 * 64
    ld = N * img_stride
    out = torch.zeros(copies * Cc, ld, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    for b in range(copies):
        shift = b - 1 if copies == 3 else 0
        _lib.check(lib.ivgpt_transpose_pad(x.data_ptr(), out.data_ptr() + b * Cc * ld * 4, N, H, W, Cc, Wp, shift, img_stride,
                                           ld, _stream()), "transpose_pad")
    return out


def zero_insert2x(dy: torch.Tensor) -> torch.Tensor:
    """[N,h,w,C] -> [N,2h,2w,C] with out[2i+1, 2j+1] = dy[i, j]: the operand of the stride-2 conv's data gradient."""
    _cuda(dy)
    _f32c(dy)
    N, h, w, Cc = dy.shape
    out = torch.empty(N, 2 * h, 2 * w, Cc, dtype=torch.float32, device=dy.device)
    _lib.check(_lib.load().ivgpt_zero_insert2x(dy.data_ptr(), out.data_ptr(), N, h, w, Cc, _stream()), "zero_insert2x")
    return out


def upsample2x_bwd(dy: torch.Tensor) -> torch.Tensor:
    _cuda(dy)
    _f32c(dy)
    N, H2, W2, Cc = dy.shape
    dx = torch.empty(N, H2 // 2, W2 // 2, Cc, dtype=torch.float32, device=dy.device)
    _lib.check(_lib.load().ivgpt_upsample2x_bwd(dy.data_ptr(), dx.data_ptr(), N, H2 // 2, W2 // 2, Cc, _stream()),
               "upsample2x_bwd")
    return dx


def silu(x: torch.Tensor, dy: Optional[torch.Tensor] = None) -> torch.Tensor:
    """forward: silu(x); with dy: dy * silu'(x)."""
    _cuda(x, dy)
    _f32c(x, dy)
    out = torch.empty_like(x)
    _lib.check(_lib.load().ivgpt_silu(x.data_ptr(), _ptr(dy), out.data_ptr(), x.numel(), _stream()), "silu")
    return out


def axpby(x: torch.Tensor, y: Optional[torch.Tensor], a: float, b: float = 0.0, out: Optional[torch.Tensor] = None):
    _cuda(x, y, out)
    _f32c(x, y, out)
    out = torch.empty_like(x) if out is None else out
    _lib.check(_lib.load().ivgpt_axpby(x.data_ptr(), _ptr(y), out.data_ptr(), float(a), float(b), x.numel(), _stream()),
               "axpby")
    return out


def reduce_mid(x: torch.Tensor, outer: int, mid: int, out: Optional[torch.Tensor] = None, accumulate: bool = False):
    """x viewed [outer, mid, inner] -> out [outer, inner] (+)= sum over mid."""
    _cuda(x, out)
    _f32c(x, out)
    inner = x.numel() // (outer * mid)
    if out is None:
        assert not accumulate
        out = torch.empty(outer, inner, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_reduce_mid(x.data_ptr(), out.data_ptr(), outer, mid, inner, int(accumulate), _stream()),
               "reduce_mid")
    return out


def nchw_to_nhwc(x: torch.Tensor, c_pad: Optional[int] = None) -> torch.Tensor:
    """[N, Cs, H, W] fp32 -> [N, H, W, c_pad] (channels beyond Cs zero)."""
    _cuda(x)
    _f32c(x)
    N, Cs, H, W = x.shape
    Cd = Cs if c_pad is None else c_pad
    y = torch.empty(N, H, W, Cd, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().ivgpt_nchw_to_nhwc(x.data_ptr(), y.data_ptr(), N, Cs, Cd, H * W, _stream()), "nchw_to_nhwc")
    return y


def vq_bwd(z: torch.Tensor, zq: torch.Tensor, dout: Optional[torch.Tensor], gloss: Optional[torch.Tensor], beta: float):
    """(dz, de_rows): straight-through + commitment gradients of the VQ layer; de_rows are scattered onto the codebook
    with embed_bwd(idx, de_rows, dE).  gloss: 0-d fp32 CUDA tensor (gradient of the returned loss) or None."""
    _cuda(z, zq, dout, gloss)
    _f32c(z, zq, dout, gloss)
    dz, de = torch.empty_like(z), torch.empty_like(z)
    _lib.check(_lib.load().ivgpt_vq_bwd(z.data_ptr(), zq.data_ptr(), _ptr(dout), _ptr(gloss), float(beta), z.numel(),
                                        dz.data_ptr(), de.data_ptr(), _stream()), "vq_bwd")
    return dz, de


# ------------------------------------------------------------------------------------------------
# device guard: every launch above goes to torch.cuda.current_stream() of the CURRENT device, and the C side keeps
# per-device state keyed by cudaGetDevice().  A tensor living on another GPU (model on cuda:1 while cuda:0 is current)
# therefore switches the current device for the duration of the call.
# ------------------------------------------------------------------------------------------------
def _first_cuda_device(args, kwargs):
    for x in args:
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    for x in kwargs.values():
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return x.device
    return None


def _device_guarded(fn):
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = _first_cuda_device(args, kwargs)
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapped


def on_device_of(t: torch.Tensor):
    """Context manager: make t's GPU the current device (no-op when it already is).  Used by the model-level entry points
    (tokenize / detokenize / generate / forward) so that descriptor-only launches (gemm_raw) follow the model's device."""
    return torch.cuda.device(t.device)


device_scoped = _device_guarded      # decorator for model-level methods (first CUDA tensor argument decides)

for _name, _fn in list(globals().items()):
    if isinstance(_fn, types.FunctionType) and not _name.startswith("_") and _fn.__module__ == __name__ and \
            _name not in ("gemm_desc", "gemm_raw", "torch_dtype", "set_pdl", "set_deterministic", "set_mma_issuers", "set_gemm_mh2", "on_device_of", "device_scoped"):
        globals()[_name] = _device_guarded(_fn)
del _name, _fn
