"""Kernel plan of the compressive tokenizer: which B200 kernels run, in what order, on what layout.

Replaces the eager module graph of reference ivideogpt/vq_model/vae.py (Encoder.forward :141-195,
Decoder.forward :298-371), conditional_vae.py (CrossAttentionBlock.forward :38-55, ConditionalEncoder.forward
:108-132, ConditionalDecoder.forward :186-212) and the diffusers blocks they call.

Layout decisions (B200-first, not a translation):
  * activations are NHWC in the compute dtype (fp32 fed to tensor cores as TF32 -- the reference's own GPU
    behaviour, cudnn.allow_tf32 -- or bf16); every 3x3 conv is ONE implicit-GEMM tcgen05 launch whose
    zero padding is the TMA out-of-bounds fill; the ResnetBlock 1x1 shortcut rides in the same launch as
    extra K blocks; residual adds / bias / SiLU live in the GEMM epilogue.
  * context features are NEVER repeated per future frame (reference compressive_vq_model.py:176-187,257-266
    materialises 14 copies); cross-attention K / V^T projections are computed once per clip and every
    (frame, head) tile of Q.K^T and P.V indexes the clip it belongs to.
  * weights are repacked once (K-major [Cout, 9*Cin(+Cin_shortcut)]) and cached per parameter version.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch

from .. import ops
from .._lib import ACT_NONE, ACT_SILU, BF16, F32
from .modules import CrossAttnParams, DecoderParams, EncoderParams, MidAttnParams, MidParams, ResnetParams


def _round_tf32(w: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest fp32 -> tf32 (10-bit mantissa) so that the tensor core's truncation is exact."""
    i = w.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


class PackedWeights:
    """Per-model cache of kernel-layout weights, keyed by (parameter identity, version, dtype)."""

    def __init__(self):
        self._cache: Dict[tuple, torch.Tensor] = {}
        self._latest: Dict[tuple, tuple] = {}

    def clear(self):
        self._cache.clear()
        self._latest.clear()

    def _key(self, tag, params, dtype):
        """Cache key.  The parameter VERSIONS are part of it, and entries of older versions of the same parameters are dropped
        when a new version is packed: during training every optimizer step bumps every version, and a cache that kept the
        stale copies would grow by one full set of packed weights per step."""
        ident = (tag, dtype) + tuple(p.data_ptr() for p in params)
        key = ident + tuple(p._version for p in params)
        old = self._latest.get(ident)
        if old is not None and old != key:
            self._cache.pop(old, None)
        self._latest[ident] = key
        return key

    def _cast(self, w: torch.Tensor, dtype):
        w = w.detach().to(torch.float32)
        if dtype == torch.float32:
            return _round_tf32(w).contiguous()
        return w.to(dtype).contiguous()

    def conv3(self, conv, dtype, shortcut=None) -> Tuple[torch.Tensor, torch.Tensor]:
        params = [conv.weight, conv.bias] + ([shortcut.weight, shortcut.bias] if shortcut is not None else [])
        key = self._key("conv3", params, dtype)
        if key not in self._cache:
            w = conv.weight.detach().float().permute(0, 2, 3, 1).reshape(conv.weight.shape[0], -1)  # [Cout, 9*Cin]
            b = conv.bias.detach().float().clone()
            if shortcut is not None:
                w = torch.cat([w, shortcut.weight.detach().float().reshape(shortcut.weight.shape[0], -1)], dim=1)
                b = b + shortcut.bias.detach().float()
            self._cache[key] = (self._cast(w, dtype), b.contiguous())
        return self._cache[key]

    def linear(self, weight, bias, dtype, tag="lin") -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        params = [weight] + ([bias] if bias is not None else [])
        key = self._key(tag, params, dtype)
        if key not in self._cache:
            w = weight.detach().float().reshape(weight.shape[0], -1)
            self._cache[key] = (self._cast(w, dtype), None if bias is None else bias.detach().float().contiguous())
        return self._cache[key]

    def rows(self, weight, bias, lo, hi, dtype, tag) -> Tuple[torch.Tensor, torch.Tensor]:
        key = self._key((tag, lo, hi), [weight, bias], dtype)
        if key not in self._cache:
            self._cache[key] = (self._cast(weight.detach().float()[lo:hi], dtype),
                                bias.detach().float()[lo:hi].contiguous())
        return self._cache[key]

    def f32(self, p) -> torch.Tensor:
        key = self._key("f32", [p], torch.float32)
        if key not in self._cache:
            self._cache[key] = p.detach().float().contiguous()
        return self._cache[key]

    def derived(self, tag, params, make):
        """Any other kernel-layout copy of `params` (training: flipped / transposed weights for the data gradients),
        rebuilt when a parameter's version changes."""
        key = self._key(tag, params, torch.float32)
        if key not in self._cache:
            self._cache[key] = make()
        return self._cache[key]

    def conv3_dgrad(self, conv, cout_pad: int = 0) -> torch.Tensor:
        """[Cin, 9*Cout] K-major weights of the conv's data gradient: dX = conv3x3(dY, this) with the taps flipped
        (Wd[ci, (a, b), co] = W[co, ci, 2-a, 2-b]); cout_pad > Cout appends zero output channels (the 3-channel conv_out)."""
        def make():
            w = conv.weight.detach().float()
            if cout_pad > w.shape[0]:
                w = torch.cat([w, w.new_zeros(cout_pad - w.shape[0], *w.shape[1:])], dim=0)
            return _round_tf32(w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous())
        return self.derived(("conv3_dgrad", cout_pad), [conv.weight], make)

    def linear_t(self, weight, lo: Optional[int] = None, hi: Optional[int] = None) -> torch.Tensor:
        """[K, N] = rows lo:hi of a Linear / 1x1-conv weight, transposed (the B operand of dX = dY @ W)."""
        def make():
            w = weight.detach().float().reshape(weight.shape[0], -1)
            if lo is not None:
                w = w[lo:hi]
            return _round_tf32(w.t().contiguous())
        return self.derived(("lin_t", lo, hi), [weight], make)

    def conv_in27(self, conv):
        key = self._key("cin", [conv.weight, conv.bias], torch.float32)
        if key not in self._cache:
            self._cache[key] = (conv.weight.detach().float().reshape(conv.weight.shape[0], 27).contiguous(),
                                conv.bias.detach().float().contiguous())
        return self._cache[key]

    def conv_out3(self, conv):
        key = self._key("cout3", [conv.weight, conv.bias], torch.float32)
        if key not in self._cache:
            w = conv.weight.detach().float().permute(0, 2, 3, 1).reshape(3, 9, -1).contiguous()  # [3][tap][C]
            self._cache[key] = (w, conv.bias.detach().float().contiguous())
        return self._cache[key]


class TokenizerPlan:
    def __init__(self, groups: int, compute_dtype: torch.dtype):
        self.groups = groups
        self.dtype = compute_dtype
        self.pw = PackedWeights()
        # GroupNorm statistics CAN be emitted by the producing conv's epilogue (ops.conv3x3(gn_groups=...)): per-row running sums
        # per group -> shared-memory partials -> fixed-order reduction once per 64-column window.  Correct, and MEASURED slower
        # again (round 2: conv family 39.9 -> 46.4 ms per cfg64 step, 127 -> 158 ms at cfg256, for a ~7 / ~15 ms statistics pass
        # saved): the epilogue is the stage the conv kernel can least afford to load.  OFF; IVGPT_FUSED_GN_STATS=1 selects it.
        self.fused_stats = groups if os.environ.get("IVGPT_FUSED_GN_STATS", "0") == "1" else 0
        # The statistics pass and the apply pass of a GroupNorm read the same tensor; running them over chunks of frames small
        # enough for the apply's read to hit the 126 MB L2 was MEASURED slower (detokenize 65 -> 75 / 96 ms with 24 / 12 MB chunks
        # at cfg64: hundreds of small launches, the host falls behind).  0 = whole tensor at once (default).
        self.gn_chunk_bytes = int(os.environ.get("IVGPT_GN_CHUNK_MB", "0")) << 20
        # GroupNorm + SiLU applied to the conv's operand tiles inside the conv kernel (transform warps of gemm_tc.cu): no
        # normalised copy of the activation in HBM.  MEASURED (profiles/r02/gn_fused_apply_ab.txt): the conv family goes from
        # 45 to 82-104 ms per cfg64 step (4 / 8 / 16 transform warps: 104 / 84 / 82) -- every input pixel is re-normalised for
        # each of the 9 taps and the extra shared-memory read + write per k-block, not the arithmetic, is the limit -- against
        # ~17 ms of GroupNorm passes saved.  OFF by default; IVGPT_FUSED_GN_APPLY=1 selects it (parity-tested either way).
        self.fused_apply = os.environ.get("IVGPT_FUSED_GN_APPLY", "0") == "1"

    # ---- building blocks --------------------------------------------------------------------------
    def _gn(self, x, norm, silu: bool, samples: Optional[int] = None, pos=None):
        n = x.shape[0] if samples is None else samples
        gamma, beta = self.pw.f32(norm.weight), self.pw.f32(norm.bias)
        posf = None if pos is None else self.pw.f32(pos)
        if getattr(x, "gn_part", None) is not None and x.gn_part[0].shape[2] == norm.num_groups:
            stats = ops.groupnorm_stats_from_parts(x, n, norm.num_groups, norm.eps)   # fused in the producing conv
            return ops.groupnorm_apply(x, stats, gamma, beta, silu, posf)
        per = x.numel() * x.element_size() // max(n, 1)                   # bytes per sample
        step = n if (self.gn_chunk_bytes <= 0 or per * n <= self.gn_chunk_bytes) else max(1, self.gn_chunk_bytes // per)
        if step >= n:
            stats = ops.groupnorm_stats(x, n, norm.num_groups, norm.eps)
            return ops.groupnorm_apply(x, stats, gamma, beta, silu, posf)
        # statistics + apply per chunk of samples: the second read of a chunk comes from L2 instead of HBM
        fps = x.shape[0] // n                                            # frames per sample (joint norms merge frames)
        y = torch.empty_like(x)
        for s0 in range(0, n, step):
            s1 = min(n, s0 + step)
            xc = x[s0 * fps: s1 * fps]
            stats = ops.groupnorm_stats(xc, s1 - s0, norm.num_groups, norm.eps)
            ops.groupnorm_apply(xc, stats, gamma, beta, silu, posf, out=y[s0 * fps: s1 * fps])
        return y

    def _gn_coeff(self, x, norm):
        """Coefficients of GroupNorm(x) for the conv that applies it on the fly: statistics pass (or the producing conv's
        epilogue partials) + a tiny per-(frame, channel) kernel."""
        n = x.shape[0]
        if getattr(x, "gn_part", None) is not None and x.gn_part[0].shape[2] == norm.num_groups:
            stats = ops.groupnorm_stats_from_parts(x, n, norm.num_groups, norm.eps)
        else:
            stats = ops.groupnorm_stats(x, n, norm.num_groups, norm.eps)
        sc, sh = ops.groupnorm_coeff(stats, self.pw.f32(norm.weight), self.pw.f32(norm.bias))
        return sc, sh, True

    def gn_silu_conv(self, x, norm, w, b, **kw):
        """conv3x3(silu(GroupNorm(x))) -- fused (default) or as two launches."""
        if self.fused_apply:
            return ops.conv3x3(x, w, b, gn_in=self._gn_coeff(x, norm), **kw)
        return ops.conv3x3(self._gn(x, norm, True), w, b, **kw)

    def resnet(self, x, r: ResnetParams):
        w1, b1 = self.pw.conv3(r.conv1, self.dtype)
        h = self.gn_silu_conv(x, r.norm1, w1, b1, gn_groups=self.fused_stats)
        if r.conv_shortcut is not None:
            w2, b2 = self.pw.conv3(r.conv2, self.dtype, shortcut=r.conv_shortcut)
            return self.gn_silu_conv(h, r.norm2, w2, b2, x2=x, gn_groups=self.fused_stats)
        w2, b2 = self.pw.conv3(r.conv2, self.dtype)
        return self.gn_silu_conv(h, r.norm2, w2, b2, residual=x, gn_groups=self.fused_stats)

    def attention(self, q_tok, kv_tok, wq, bq, wk, bk, wv, bv, heads: int, frames_per_clip: int):
        """q_tok [F, Lq, C], kv_tok [B, Lkv, C] (F = B*frames_per_clip) -> O [F*Lq, C] (before out-proj)."""
        F_, Lq, Cc = q_tok.shape
        B, Lkv, _ = kv_tok.shape
        dh = Cc // heads
        dt = self.dtype
        code = BF16 if dt == torch.bfloat16 else F32
        dev = q_tok.device
        qp = ops.gemm(q_tok.view(F_ * Lq, Cc), wq, bq)                      # [F*Lq, C]
        kp = ops.gemm(kv_tok.view(B * Lkv, Cc), wk, bk)                     # [B*Lkv, C]
        vt = torch.empty(B, Cc, Lkv, dtype=dt, device=dev)                 # V^T per clip
        ops.gemm_raw(ops.gemm_desc(
            dtype=code, a=wv.data_ptr(), lda=Cc, a_rows=Cc, a_cols=Cc, a_batches=1,
            b=kv_tok.data_ptr(), ldb=Cc, b_bstride=Lkv * Cc, b_rows=Lkv, b_cols=Cc, b_batches=B,
            M=Cc, N=Lkv, K=Cc, batch=B, heads=1, a_bsel=0, b_bsel=2, o_bsel=2,
            out=vt.data_ptr(), ldo=Lkv, out_bstride=Cc * Lkv, out_dtype=code,
            bias=bv.data_ptr(), bias_along_m=1))
        s = torch.empty(F_ * heads, Lq, Lkv, dtype=torch.float32, device=dev)
        ops.gemm_raw(ops.gemm_desc(
            dtype=code, a=qp.data_ptr(), lda=Cc, a_bstride=Lq * Cc, a_rows=Lq, a_cols=Cc, a_batches=F_,
            b=kp.data_ptr(), ldb=Cc, b_bstride=Lkv * Cc, b_rows=Lkv, b_cols=Cc, b_batches=B,
            M=Lq, N=Lkv, K=dh, batch=F_ * heads, heads=heads, a_bsel=1, a_bdiv=1, b_bsel=1, b_bdiv=frames_per_clip,
            o_bsel=2, a_khead=dh, b_khead=dh,
            out=s.data_ptr(), ldo=Lkv, out_bstride=Lq * Lkv, out_dtype=F32, alpha=1.0 / math.sqrt(dh)))
        p = torch.empty(F_ * heads, Lq, Lkv, dtype=dt, device=dev)
        ops.softmax(s, p, F_ * heads * Lq, Lq, Lkv, Lkv, Lkv, False)
        o = torch.empty(F_ * Lq, Cc, dtype=dt, device=dev)
        ops.gemm_raw(ops.gemm_desc(
            dtype=code, a=p.data_ptr(), lda=Lkv, a_bstride=Lq * Lkv, a_rows=Lq, a_cols=Lkv, a_batches=F_ * heads,
            b=vt.data_ptr(), ldb=Lkv, b_bstride=Cc * Lkv, b_rows=Cc, b_cols=Lkv, b_batches=B,
            M=Lq, N=dh, K=Lkv, batch=F_ * heads, heads=heads, a_bsel=2, b_bsel=1, b_bdiv=frames_per_clip,
            o_bsel=1, b_nhead=dh, o_nhead=dh,
            out=o.data_ptr(), ldo=Cc, out_bstride=Lq * Cc, out_dtype=code))
        return o

    def cross_attention(self, z, ctx_feat, blk: CrossAttnParams, clips: int):
        """z [F,H,W,C] frame features; ctx_feat [clips*t,H,W,C] context features of the same stage."""
        F_, H, W, Cc = z.shape
        t = ctx_feat.shape[0] // clips
        assert blk.kv_frames == t, f"cross-attention built for {blk.kv_frames} context frames, got {t}"
        kv = self._gn(ctx_feat, blk.kv_norm, False, samples=clips, pos=blk.kv_pos_emb)   # joint norm over t frames
        q = self._gn(z, blk.q_norm, False, pos=blk.q_pos_emb)
        dt = self.dtype
        wq, bq = self.pw.rows(blk.att.in_proj_weight, blk.att.in_proj_bias, 0, Cc, dt, "mha_q")
        wk, bk = self.pw.rows(blk.att.in_proj_weight, blk.att.in_proj_bias, Cc, 2 * Cc, dt, "mha_k")
        wv, bv = self.pw.rows(blk.att.in_proj_weight, blk.att.in_proj_bias, 2 * Cc, 3 * Cc, dt, "mha_v")
        o = self.attention(q.view(F_, H * W, Cc), kv.view(clips, t * H * W, Cc), wq, bq, wk, bk, wv, bv,
                           blk.heads, F_ // clips)
        wo, bo = self.pw.linear(blk.att.out_proj.weight, blk.att.out_proj.bias, dt)
        out = ops.gemm(o, wo, bo, residual=z.view(F_ * H * W, Cc), act=ACT_SILU)         # silu(z + attn)
        return out.view(F_, H, W, Cc)

    def mid_attention(self, x, a: MidAttnParams):
        F_, H, W, Cc = x.shape
        tok = self._gn(x, a.group_norm, False).view(F_, H * W, Cc)
        dt = self.dtype
        wq, bq = self.pw.linear(a.to_q.weight, a.to_q.bias, dt)
        wk, bk = self.pw.linear(a.to_k.weight, a.to_k.bias, dt)
        wv, bv = self.pw.linear(a.to_v.weight, a.to_v.bias, dt)
        o = self.attention(tok, tok, wq, bq, wk, bk, wv, bv, 1, 1)
        wo, bo = self.pw.linear(a.to_out[0].weight, a.to_out[0].bias, dt)
        return ops.gemm(o, wo, bo, residual=x.view(F_ * H * W, Cc)).view(F_, H, W, Cc)

    def mid(self, x, m: MidParams):
        x = self.resnet(x, m.resnets[0])
        if m.has_attention:
            x = self.mid_attention(x, m.attentions[0])
        return self.resnet(x, m.resnets[1])

    # ---- encoders ---------------------------------------------------------------------------------
    def encode(self, clips_px, enc: EncoderParams, frame_offset: int, frames_per_clip: int,
               ctx_feats: Optional[List[torch.Tensor]] = None, want_features: bool = False):
        """clips_px [B,T,3,H,W] fp32.  Returns latent [F,16,16,latent] (+ stage features when asked)."""
        B = clips_px.shape[0]
        w27, b27 = self.pw.conv_in27(enc.conv_in)
        x = ops.conv_in(clips_px, w27, b27, self.dtype, frame_offset, frames_per_clip)
        feats = [x]
        k = 0
        for i, stage in enumerate(enc.down_blocks):
            for r in stage.resnets:
                x = self.resnet(x, r)
            if stage.downsamplers is not None:
                wd, bd = self.pw.conv3(stage.downsamplers[0].conv, self.dtype)
                x = ops.conv3x3(x, wd, bd, stride=2, gn_groups=self.fused_stats)
            if ctx_feats is not None and x.shape[2] <= enc.max_att_resolution:
                x = self.cross_attention(x, ctx_feats[i + 1], enc.cross_att_blocks[k], B)
                k += 1
            feats.append(x)
        x = self.mid(x, enc.mid_block)
        feats.append(x)
        wo, bo = self.pw.conv3(enc.conv_out, self.dtype)
        out = self.gn_silu_conv(x, enc.conv_norm_out, wo, bo)
        return (out, feats) if want_features else out

    # ---- decoders ---------------------------------------------------------------------------------
    def decode(self, latent, dec: DecoderParams, out_clips, frame_offset: int, frames_per_clip: int,
               ctx_feats: Optional[List[torch.Tensor]] = None, want_features: bool = False):
        """latent [F,16,16,latent] NHWC -> frames written into out_clips [B,T,3,H,W]."""
        B = out_clips.shape[0]
        wi, bi = self.pw.conv3(dec.conv_in, self.dtype)
        x = ops.conv3x3(latent, wi, bi, gn_groups=self.fused_stats)
        feats = [x]
        x = self.mid(x, dec.mid_block)
        feats.append(x)
        if ctx_feats is not None:
            x = self.cross_attention(x, ctx_feats[1], dec.cross_att_blocks[0], B)
        for i, stage in enumerate(dec.up_blocks):
            for r in stage.resnets:
                x = self.resnet(x, r)
            if stage.upsamplers is not None:
                x = ops.upsample2x(x)
                wu, bu = self.pw.conv3(stage.upsamplers[0].conv, self.dtype)
                x = ops.conv3x3(x, wu, bu, gn_groups=self.fused_stats)
            if ctx_feats is not None and x.shape[2] <= dec.max_att_resolution:
                x = self.cross_attention(x, ctx_feats[i + 2], dec.cross_att_blocks[i + 1], B)
            feats.append(x)
        if getattr(x, "gn_part", None) is not None:
            stats = ops.groupnorm_stats_from_parts(x, x.shape[0], dec.conv_norm_out.num_groups, dec.conv_norm_out.eps)
        else:
            stats = ops.groupnorm_stats(x, x.shape[0], dec.conv_norm_out.num_groups, dec.conv_norm_out.eps)
        w3, b3 = self.pw.conv_out3(dec.conv_out)
        ops.conv_out3(x, stats, self.pw.f32(dec.conv_norm_out.weight), self.pw.f32(dec.conv_norm_out.bias), w3, b3,
                      out_clips, frame_offset, frames_per_clip)
        return feats if want_features else None
