"""Kernel plan for the Llama-style autoregressive transformer over frame tokens.

Replaces transformers' LlamaModel / LlamaAttention / LlamaMLP / GenerationMixin arithmetic that the reference
reaches through `model.generate(...)` (inference/predict.py:64-69, ivideogpt/transformer/action_model.py:86-110)
and `model(input_ids=, labels=)` (train_gpt.py:792).

B200-first layout:
  * weights are packed once per parameter version: [q;k;v] fused to one [3h,h] operand, gate/up interleaved
    row-wise into [2*inter,h] so SwiGLU is a GEMM epilogue, all in the compute dtype (bf16, or fp32->TF32);
  * the residual stream stays fp32; every projection is one tcgen05 GEMM launch with bias-free epilogues that
    add the residual in place;
  * static KV cache, K as [B,heads,Lmax,64] and V transposed [B,heads,64,Lmax]: prefill attention is
    Q.K^T -> causal softmax -> P.V^T^T on the tensor cores straight out of the cache, decode attention is a
    single HBM-streaming kernel over the same buffers;
  * one decode step (embed -> 12/24 layers -> lm_head -> sample -> append) is captured once into a CUDA graph
    whose kernels read the current position from device memory, and replayed max_new_tokens-1 times.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import os

import torch

from .. import ops
from .._lib import ACT_SWIGLU, BF16, F32


def _round_tf32(w: torch.Tensor) -> torch.Tensor:
    i = w.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _pack(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    w = w.detach().float()
    return _round_tf32(w).contiguous() if dtype == torch.float32 else w.to(dtype).contiguous()


class LlamaWeights:
    """Kernel-layout copy of an HF LlamaForCausalLM's parameters (cached per parameter version)."""

    def __init__(self, hf_model, dtype: torch.dtype):
        cfg = hf_model.config
        self.dtype = dtype
        self.hidden = cfg.hidden_size
        self.inter = cfg.intermediate_size
        self.heads = cfg.num_attention_heads
        self.layers_n = cfg.num_hidden_layers
        self.vocab = cfg.vocab_size
        self.eps = float(cfg.rms_norm_eps)
        if cfg.num_key_value_heads != cfg.num_attention_heads:
            raise NotImplementedError("grouped-query attention is not used by iVideoGPT configs and not implemented")
        if self.hidden // self.heads != 64:
            raise NotImplementedError("head_dim must be 64 (both iVideoGPT Llama configs)")
        if getattr(cfg, "attention_bias", False) or getattr(cfg, "mlp_bias", False):
            raise NotImplementedError("attention/mlp biases are not part of the iVideoGPT Llama configs")
        m = hf_model.model
        self.embed = m.embed_tokens.weight.detach().float().contiguous()
        self.layers = []
        for lyr in m.layers:
            a, mlp = lyr.self_attn, lyr.mlp
            wqkv = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], dim=0)
            wgu = torch.stack([mlp.gate_proj.weight, mlp.up_proj.weight], dim=1).reshape(2 * self.inter, self.hidden)
            self.layers.append(dict(
                wqkv=_pack(wqkv, dtype), wo=_pack(a.o_proj.weight, dtype), wgu=_pack(wgu, dtype),
                wd=_pack(mlp.down_proj.weight, dtype),
                n1=lyr.input_layernorm.weight.detach().float().contiguous(),
                n2=lyr.post_attention_layernorm.weight.detach().float().contiguous()))
        self.norm = m.norm.weight.detach().float().contiguous()
        self.lm_head = _pack(hf_model.lm_head.weight, dtype)
        import weakref
        self._src = weakref.ref(hf_model)          # fp32 masters, for the norm-folded copies of the decode megakernel
        self._folded = None
        # RoPE tables exactly as LlamaRotaryEmbedding computes them (fp32, theta from config)
        theta = float(getattr(cfg, "rope_theta", None) or (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0))
        dev = self.embed.device
        inv_freq = 1.0 / (theta ** (torch.arange(0, 64, 2, dtype=torch.int64, device=dev).float() / 64))
        pos = torch.arange(cfg.max_position_embeddings, device=dev, dtype=torch.float32)
        freqs = pos[:, None] * inv_freq[None, :]
        self.cos, self.sin = freqs.cos().contiguous(), freqs.sin().contiguous()
        self.max_pos = cfg.max_position_embeddings

    def folded(self):
        """RMSNorm weights folded into the COLUMNS of the matrices that consume the normalised activations (decode megakernel
        with fused norms: xn @ W^T = rstd * (x @ (W (.) g)^T)): per layer wqkv (.) n1 and wgu (.) n2, lm_head (.) final norm;
        computed from the fp32 parameters, rounded once to the compute dtype.  Built on first use, cached per weight version."""
        if self._folded is None:
            hf = self._src()
            if hf is None:
                raise RuntimeError("the model these kernel-layout weights were packed from no longer exists")
            layers = []
            for lyr in hf.model.layers:
                a, mlp = lyr.self_attn, lyr.mlp
                n1 = lyr.input_layernorm.weight.detach().float()[None, :]
                n2 = lyr.post_attention_layernorm.weight.detach().float()[None, :]
                wqkv = torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], dim=0).detach().float() * n1
                wgu = torch.stack([mlp.gate_proj.weight, mlp.up_proj.weight], dim=1).reshape(2 * self.inter, self.hidden).detach().float() * n2
                layers.append(dict(wqkv=_pack(wqkv, self.dtype), wgu=_pack(wgu, self.dtype)))
            lm = hf.lm_head.weight.detach().float() * hf.model.norm.weight.detach().float()[None, :]
            self._folded = dict(layers=layers, lm_head=_pack(lm, self.dtype))
        return self._folded

    @staticmethod
    def signature(hf_model, dtype):
        return (dtype,) + tuple((p.data_ptr(), p._version) for p in hf_model.parameters())


class LlamaEngine:
    def __init__(self, weights: LlamaWeights):
        self.w = weights
        self.dtype = weights.dtype
        self.code = BF16 if self.dtype == torch.bfloat16 else F32
        self._buf: Dict[tuple, torch.Tensor] = {}
        self._graphs: Dict[tuple, tuple] = {}
        # bf16 prefill attention: fused kernel (default) or the round-1 materialised path (scores GEMM -> softmax -> PV GEMM)
        self.use_flash = os.environ.get("IVGPT_FLASH_ATTN", "1") == "1"

    # ---- buffers ------------------------------------------------------------------------------------
    def buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._buf.get(key)
        if t is None:
            t = torch.empty(*shape, dtype=dtype, device=self.w.embed.device)
            self._buf[key] = t
        return t

    def kv_cache(self, B, Lmax):
        w = self.w
        k = self.buf("kcache", (w.layers_n, B, w.heads, Lmax, 64), self.dtype)
        v = self.buf("vcache_t", (w.layers_n, B, w.heads, 64, Lmax), self.dtype)
        return k, v

    def v_rows(self, B, Lmax):
        """Second V cache in K's [Lmax][64] layout: written by the prefill's rope_kv, streamed by the decode megakernel."""
        w = self.w
        return self.buf("vcache_rows", (w.layers_n, B, w.heads, Lmax, 64), self.dtype)

    # ---- one transformer layer over M = B*Lq rows -----------------------------------------------------
    def _layer(self, li, x, B, Lq, kc, vc, Lmax, pos0, dpos, prefill: bool):
        w, dt, code = self.w, self.dtype, self.code
        lw = w.layers[li]
        M = B * Lq
        h, H = w.hidden, w.heads
        xn = self.buf("xn", (M, h), dt)
        ops.rmsnorm(x, lw["n1"], xn, M, w.eps)
        qkv = self.buf("qkv", (M, 3 * h), dt)
        ops.gemm(xn, lw["wqkv"], out=qkv)
        q = self.buf("q", (B, H, Lq, 64), dt)
        vr = self.v_rows(B, Lmax)[li] if (prefill and dt == torch.bfloat16) else None
        ops.rope_kv(qkv, q, kc[li], vc[li], B, Lq, H, Lmax, pos0, dpos, w.cos, w.sin, v_rows=vr)
        ao = self.buf("attn_out", (M, h), dt)
        if prefill and dt == torch.bfloat16 and self.use_flash:
            # fused QK^T -> online softmax -> PV on tcgen05, S in TMEM, P through shared memory (csrc/flash_attn.cu)
            ops.flash_attn(q, kc[li], vc[li], ao, B, H, Lq, pos0 + Lq, Lmax * 64, 64 * Lmax, Lmax, causal=True, scale=0.125)
        elif prefill:
            Lk = pos0 + Lq
            ld = (Lk + 7) // 8 * 8
            s = self.buf("scores", (B * H, Lq, ld), torch.float32)
            ops.gemm_raw(ops.gemm_desc(
                dtype=code, a=q.data_ptr(), lda=64, a_bstride=Lq * 64, a_rows=Lq, a_cols=64, a_batches=B * H,
                b=kc[li].data_ptr(), ldb=64, b_bstride=Lmax * 64, b_rows=Lk, b_cols=64, b_batches=B * H,
                M=Lq, N=Lk, K=64, batch=B * H, heads=1, a_bsel=2, b_bsel=2, o_bsel=2, causal_skip=1 if pos0 == 0 else 0,
                out=s.data_ptr(), ldo=ld, out_bstride=Lq * ld, out_dtype=F32, alpha=0.125))
            p = self.buf("probs", (B * H, Lq, ld), dt)
            ops.softmax(s, p, B * H * Lq, Lq, Lk, ld, ld, True, pos0)
            ops.gemm_raw(ops.gemm_desc(
                dtype=code, a=p.data_ptr(), lda=ld, a_bstride=Lq * ld, a_rows=Lq, a_cols=Lk, a_batches=B * H,
                b=vc[li].data_ptr(), ldb=Lmax, b_bstride=64 * Lmax, b_rows=64, b_cols=Lk, b_batches=B * H,
                M=Lq, N=64, K=Lk, batch=B * H, heads=H, a_bsel=2, b_bsel=2, o_bsel=1, o_nhead=64,
                out=ao.data_ptr(), ldo=h, out_bstride=Lq * h, out_dtype=code))
        else:
            ops.decode_attn(q, kc[li], vc[li], ao, B, H, Lmax, pos0 + 1, dpos, 0.125)
        ops.gemm(ao, lw["wo"], residual=x, out=x)                      # x += attn @ Wo^T   (fp32, in place)
        ops.rmsnorm(x, lw["n2"], xn, M, w.eps)
        act = self.buf("act", (M, w.inter), dt)
        ops.gemm(xn, lw["wgu"], act=ACT_SWIGLU, out=act)               # silu(gate) * up
        ops.gemm(act, lw["wd"], residual=x, out=x)                     # x += act @ Wd^T

    # ---- prefill ------------------------------------------------------------------------------------------
    def prefill(self, B, L, Lmax, ids: Optional[torch.Tensor], embeds: Optional[torch.Tensor], all_logits: bool,
                logits_out: Optional[torch.Tensor] = None, want_hidden: bool = False, hidden_only: bool = False):
        """Runs L prompt positions, fills the KV cache.  Returns logits: [B, V] of the last position, or
        [B, L, Vpad] of every position when all_logits."""
        w = self.w
        M = B * L
        h = w.hidden
        if Lmax > w.max_pos:
            raise ValueError(f"sequence length {Lmax} exceeds max_position_embeddings {w.max_pos}")
        x = self.buf("x", (M, h), torch.float32)
        if embeds is not None:
            assert embeds.shape == (B, L, h)
            x.copy_(embeds.reshape(M, h).to(torch.float32))
        else:
            assert ids.dtype == torch.int64 and ids.is_contiguous()
            ops.embed(ids, ids.stride(0), L, None, w.embed, x, M)
        kc, vc = self.kv_cache(B, Lmax)
        for li in range(w.layers_n):
            self._layer(li, x, B, L, kc, vc, Lmax, 0, None, True)
        xn = self.buf("xn", (M, h), self.dtype)
        ops.rmsnorm(x, w.norm, xn, M, w.eps)
        V = w.vocab
        if hidden_only:                                  # final-norm states of every position, no lm_head
            return None, xn.view(B, L, h)
        if all_logits:
            vpad = (V + 3) // 4 * 4
            logits = logits_out if logits_out is not None else torch.empty(B, L, vpad, dtype=torch.float32,
                                                                           device=x.device)
            ops.gemm(xn, w.lm_head, out=logits.view(M, vpad)[:, :V])
            hidden = xn.view(B, L, h) if want_hidden else None
            return logits, hidden
        logits = self.buf("logits", (B, (V + 3) // 4 * 4), torch.float32)
        last = xn.view(B, L, h)[:, L - 1, :]                         # strided view: lda = L*h, no copy
        ops.gemm(last, w.lm_head, out=logits[:, :V])
        return logits, None

    # ---- one decode step, position read from device memory ----------------------------------------------------
    def _decode_step(self, B, Lmax, tokens, dpos, sample_cfg, dseed=None, slot=None):
        w = self.w
        h = w.hidden
        x = self.buf("xd", (B, h), torch.float32)
        ops.embed(tokens, tokens.stride(0), 1, dpos, w.embed, x, B)
        if slot is not None and slot[3] is not None:
            ops.slot_embed_add(x, slot[3], dpos, B, h, slot[0], slot[1])
        kc, vc = self.kv_cache(B, Lmax)
        for li in range(w.layers_n):
            self._layer_decode(li, x, B, kc, vc, Lmax, dpos)
        xn = self.buf("xnd", (B, h), self.dtype)
        ops.rmsnorm(x, w.norm, xn, B, w.eps)
        V = w.vocab
        logits = self.buf("logits", (B, (V + 3) // 4 * 4), torch.float32)
        ops.gemm(xn, w.lm_head, out=logits[:, :V])
        self._sample(logits, B, tokens, dpos, sample_cfg, 0, dseed)
        if slot is not None:
            ops.slot_force(tokens, dpos, B, slot[0], slot[1], slot[2])
        ops.incr(dpos, 1)

    def _layer_decode(self, li, x, B, kc, vc, Lmax, dpos):
        w, dt = self.w, self.dtype
        lw = w.layers[li]
        h, H = w.hidden, w.heads
        xn = self.buf("xnd", (B, h), dt)
        ops.rmsnorm(x, lw["n1"], xn, B, w.eps)
        qkv = self.buf("qkvd", (B, 3 * h), dt)
        ops.gemm(xn, lw["wqkv"], out=qkv)
        ao = self.buf("aod", (B, h), dt)
        ops.decode_attn_fused(qkv, kc[li], vc[li], ao, B, H, Lmax, 0, dpos, w.cos, w.sin, 0.125)
        ops.gemm(ao, lw["wo"], residual=x, out=x)
        ops.rmsnorm(x, lw["n2"], xn, B, w.eps)
        act = self.buf("actd", (B, w.inter), dt)
        ops.gemm(xn, lw["wgu"], act=ACT_SWIGLU, out=act)
        ops.gemm(act, lw["wd"], residual=x, out=x)

    def _sample(self, logits, B, tokens, dpos, sample_cfg, out_offset, dseed=None):
        """Writes the next token of every row to tokens[b, (*dpos + 1 if dpos else 0) + out_offset]."""
        V = self.w.vocab
        ld = logits.stride(0)
        if sample_cfg is None:
            ops.argmax(logits, ld, B, V, tokens, tokens.stride(0), dpos, out_offset)
        else:
            k, temp = sample_cfg
            ops.topk_sample(logits, ld, B, V, k, temp, 0, 0, tokens, tokens.stride(0), dpos, out_offset, dseed)

    # ---- persistent decode megakernel (bf16, B <= 64 with these widths) --------------------------------------
    def mega_mode(self) -> int:
        """GEMM phases of the megakernel: 0 = activation-stationary (default, fastest measured), 1 = weight-stationary (64 weight
        rows per MMA: 2.4x fewer tcgen05.mma issues but 285 vs 230 ms per rollout, profiles/r01/mega_build_variants_ab2.txt)."""
        return int(getattr(self, "mega_gemm_mode", int(os.environ.get("IVGPT_MEGA_GEMM", "0"))))

    def mega_supported(self, B: int, Lmax: int) -> bool:
        w = self.w
        if self.dtype != torch.bfloat16 or B < 1 or B > 128:
            return False
        common = (w.vocab + 256) * 4 <= 96 * 1024 and (Lmax + 72) * 4 * 8 <= 30 * 1024
        if self.mega_mode() == 1:
            if B > 64:        # the experimental weight-stationary mode is validated for B <= 64 only (a B = 100 run showed
                return False  # 7 % logit error on some rows, tests log of round 1); larger batches take the CUDA-graph path
            q_s, o_s, d_s = self._mega_splits64()
            a_rows = (B + 7) // 8 * 8
            a_bytes = max(a_rows * max(w.hidden, w.inter // d_s) * 2, 128 * 1024)
            return common and w.hidden % 64 == 0 and w.inter % 64 == 0 and a_bytes + 2 * 16 * 1024 <= 192 * 1024
        a_rows = 64 if B <= 64 else 128
        o_s, d_s = self._mega_splits()
        if o_s is None:
            return False
        kd = w.inter // d_s
        slab = max(16 * w.hidden * 2, self._mega_down()[0] * kd * 2)
        return (a_rows * w.hidden * 2 <= 128 * 1024 and a_rows * kd * 2 <= 128 * 1024 and
                a_rows * max(w.hidden, kd) * 2 + 2 * slab <= 192 * 1024 and common)

    def _mega_splits(self):
        w = self.w
        o_s = next((s for s in (3, 2, 4, 1) if w.hidden % (64 * s) == 0), None)
        if o_s is None or self._mega_down() is None:
            return None, None
        return o_s, self._mega_down()[1]

    def _mega_down(self, sms: int = 148):
        """gemm_mode 0 down projection: (tile width, split-K).  Default (16, first split with K <= 1024); `mega_down = (bn, s)` /
        IVGPT_MEGA_DOWN=bn:s selects wider tiles with more splits (same number of work items, fewer tcgen05.mma issues per
        CTA, more fp32 partials for the norm phase to add)."""
        w = self.w
        d_s_default = next((s for s in range(1, 9) if w.inter % (64 * s) == 0 and w.inter // s <= 1024), None)
        if d_s_default is None:
            return None
        ov = getattr(self, "mega_down", None)
        if ov is None and os.environ.get("IVGPT_MEGA_DOWN"):
            ov = tuple(int(v) for v in os.environ["IVGPT_MEGA_DOWN"].split(":"))
        if ov:
            bn, s = int(ov[0]), int(ov[1])
            assert bn % 16 == 0 and 16 <= bn <= 64 and 1 <= s <= 8 and w.inter % (64 * s) == 0 and w.inter // s <= 1024, ov
            return bn, s
        # Default: when 16-row items already fill one round over the SMs, 32-row tiles with twice the splits keep the item
        # count and halve the tcgen05.mma issues per CTA (138 M model: 24 tiles x 6 splits, K = 512: 32 MMAs instead of 64;
        # same-box A/B 213.7 -> 210.7 ms per rollout, logits equal to the (16, 3) path to 1e-4, tools/diag_down.py).
        s2 = 2 * d_s_default
        if (s2 <= 8 and w.inter % (64 * s2) == 0 and -(-w.hidden // 16) * d_s_default >= 128
                and -(-w.hidden // 32) * s2 <= sms and 32 * (w.inter // s2) * 2 <= 32 * 1024):
            return 32, s2
        return 16, d_s_default

    def _mega_splits64(self, sms: int = 148):
        """Split-K factors of the weight-stationary phases: the largest s <= 12 with (rows / 64) * s <= #SMs work items
        (one round over the SMs) and K a multiple of 64 * s."""
        w = self.w

        def pick(rows, K, cap, override):
            tiles = (rows + 63) // 64
            if override:
                assert K % (64 * override) == 0 and override <= cap, (rows, K, override)
                return int(override)
            best = 1
            for s in range(1, cap + 1):
                if K % (64 * s) == 0 and tiles * s <= sms:
                    best = s
            return best
        ov = getattr(self, "mega_splits_override", None) or tuple(
            int(v) for v in os.environ.get("IVGPT_MEGA_SPLITS", "0:0:0").split(":"))
        return (pick(3 * w.hidden, w.hidden, 4, ov[0]), pick(w.hidden, w.hidden, 12, ov[1]),
                pick(w.hidden, w.inter, 12, ov[2]))

    def _mega_bn_wide(self, B: int, sms: int = 148) -> int:
        """gemm_mode 0: weight rows per work item of the gate/up and lm_head phases (a multiple of 16): as wide as
        what fits in shared memory next to the K = hidden activation slab (at most 64)."""
        ov = int(getattr(self, "mega_bn_wide", int(os.environ.get("IVGPT_MEGA_BNWIDE", "0"))))
        w = self.w
        a_rows = 64 if B <= 64 else 128
        fit = min(64, (192 * 1024 - a_rows * w.hidden * 2) // (w.hidden * 2) // 16 * 16)
        if ov:
            assert ov % 16 == 0 and 16 <= ov <= fit, f"mega_bn_wide {ov} does not fit (max {fit})"
            return ov
        if -(-2 * w.inter // 16) <= sms:      # gate/up already is one round of 16-row items: nothing to gain
            return 16
        # the widest that fits: same-box A/B at B=64, hidden 768 (profiles/r01/mega_build_variants_ab4_wide_tiles.txt):
        # 16 -> 228.2 ms per rollout, 32 -> 220.9, 48 -> 215.5, 64 -> 213.7 (96 gate/up items, 2 lm_head rounds)
        return max(16, fit)

    def _mega_tables(self, bn_wide: int = 16, bn_down: int = 16):
        """Packed weight copies (swizzled slab images: 16 rows per work item in mode 0, ivgpt_mega_pack_weight; 64 rows in
        the weight-stationary mode 1, ivgpt_mega_pack_weight64) and the device-resident array of per-layer pointer
        records, built once per engine and mode."""
        mode = self.mega_mode()
        cache = getattr(self, "_mega_dev", None)
        if cache is None:
            cache = self._mega_dev = {}
        key = (mode, bn_wide if mode == 0 else 0, bn_down if mode == 0 else 0)
        if key in cache:
            return cache[key]
        import ctypes as C
        from .. import _lib
        lib = _lib.load()
        w = self.w
        dev = w.embed.device
        keep = []

        def pack(t, swiglu_pairs=0, bn=16):
            rows, cols = t.shape
            assert t.dtype == torch.bfloat16 and t.is_contiguous()
            if mode == 0 and bn != 16:
                out = torch.empty(int(lib.ivgpt_mega_packed_elems_bn(rows, cols, bn)), dtype=torch.bfloat16, device=dev)
                _lib.check(lib.ivgpt_mega_pack_weight_bn(t.data_ptr(), out.data_ptr(), rows, cols, bn, ops._stream()),
                           "mega_pack_weight_bn")
            elif mode == 1:
                out = torch.empty(int(lib.ivgpt_mega_packed_elems64(rows, cols)), dtype=torch.bfloat16, device=dev)
                _lib.check(lib.ivgpt_mega_pack_weight64(t.data_ptr(), out.data_ptr(), rows, cols, swiglu_pairs, ops._stream()),
                           "mega_pack_weight64")
            else:
                out = torch.empty(int(lib.ivgpt_mega_packed_elems(rows, cols)), dtype=torch.bfloat16, device=dev)
                _lib.check(lib.ivgpt_mega_pack_weight(t.data_ptr(), out.data_ptr(), rows, cols, ops._stream()),
                           "mega_pack_weight")
            keep.append(out)
            return out.data_ptr()

        nbytes = lib.ivgpt_mega_layer_bytes()
        host = (C.c_uint8 * (nbytes * w.layers_n + 64))()
        base_al = (C.addressof(host) + 63) // 64 * 64
        # fused norms (gemm_mode 0 of this build): the matrices fed by a normalised activation carry the norm weight
        fold = w.folded() if (mode == 0 and int(lib.ivgpt_mega_fused_norm()) == 1) else None
        for i, lw in enumerate(w.layers):
            wqkv = fold["layers"][i]["wqkv"] if fold else lw["wqkv"]
            wgu = fold["layers"][i]["wgu"] if fold else lw["wgu"]
            _lib.check(lib.ivgpt_mega_fill_layer(base_al + i * nbytes, pack(wqkv), pack(lw["wo"]), pack(wgu, 1, bn_wide),
                                                 pack(lw["wd"], 0, bn_down), lw["n1"].data_ptr(), lw["n2"].data_ptr()),
                       "mega_fill_layer")
        lm_head = pack(fold["lm_head"] if fold else w.lm_head, 0, bn_wide)
        if fold:
            w._folded = None                       # the packed copies are what the kernel reads; drop the intermediates
        raw = bytes((C.c_uint8 * (nbytes * w.layers_n)).from_address(base_al))
        tab = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        cache[key] = (tab, lm_head, keep)
        return cache[key]

    def _decode_mega(self, B, Lmax, tokens, dpos, sample_cfg, dseed, steps, slot=None):
        import ctypes as C
        from .. import _lib
        w = self.w
        h = w.hidden
        mode = self.mega_mode()
        bn_wide = self._mega_bn_wide(B) if mode == 0 else 16
        bn_down = self._mega_down()[0] if mode == 0 else 16
        dev_tab, lm_head_packed, _ = self._mega_tables(bn_wide, bn_down)
        if mode == 1:
            q_s, o_s, d_s = self._mega_splits64()
        else:
            q_s = 1
            o_s, d_s = self._mega_splits()
        kc, vc = self.kv_cache(B, Lmax)
        logits = self.buf("logits", (B, (w.vocab + 3) // 4 * 4), torch.float32)
        # ints: [0] barrier, [1] error, [512, 640) o-proj / down-proj tile counters, [1024, 2048) attention part counters
        sync = self.buf("mega_sync", (2048,), torch.int32)
        sync.zero_()
        d = _lib.MegaDesc()
        d.B, d.hidden, d.inter, d.heads, d.layers, d.vocab, d.Lmax, d.steps = B, h, w.inter, w.heads, w.layers_n, w.vocab, Lmax, steps
        d.o_splits, d.d_splits = o_s, d_s
        d.eps = w.eps
        d.x = self.buf("xd", (B, h), torch.float32).data_ptr()
        # rows of the swizzled activation images: the MMA's N in the weight-stationary mode, 64 / 128 MMA rows in mode 0
        a_rows = (B + 7) // 8 * 8 if mode == 1 else (64 if B <= 64 else 128)
        d.gemm_mode, d.qkv_splits, d.a_rows, d.bn_wide, d.bn_down = mode, q_s, a_rows, bn_wide, bn_down
        if mode == 1:
            d.qkvp = self.buf("mega_qkvp", (q_s, B, 3 * h), torch.float32).data_ptr()
        d.xn = self.buf("mega_xn", (a_rows, h), self.dtype).data_ptr()
        d.qkv = self.buf("qkvd", (B, 3 * h), self.dtype).data_ptr()
        d.ao = self.buf("mega_ao", (a_rows, h), self.dtype).data_ptr()
        d.act = self.buf("mega_act", (a_rows, w.inter), self.dtype).data_ptr()
        d.part = self.buf("mega_part", (max(o_s, d_s), B, h), torch.float32).data_ptr()
        d.logits = logits.data_ptr(); d.ldl = logits.stride(0)
        d.kcache = kc.data_ptr(); d.vcache = vc.data_ptr(); d.vrows = self.v_rows(B, Lmax).data_ptr()
        d.embed = w.embed.data_ptr(); d.norm_f = w.norm.data_ptr(); d.cos_tab = w.cos.data_ptr(); d.sin_tab = w.sin.data_ptr()
        d.tokens = tokens.data_ptr(); d.tok_stride = tokens.stride(0)
        d.dpos = dpos.data_ptr()
        if sample_cfg is None:
            d.do_sample, d.topk, d.inv_temp = 0, 1, 1.0
        else:
            d.do_sample, d.topk, d.inv_temp = 1, sample_cfg[0], 1.0 / sample_cfg[1]
        d.dseed = dseed.data_ptr()
        d.barrier = sync.data_ptr(); d.error = sync.data_ptr() + 4
        d.layers_dev = dev_tab.data_ptr(); d.lm_head_packed = lm_head_packed
        d.attn_part = self.buf("mega_attn_part", (8 * 160 * 72,), torch.float32).data_ptr()   # <= 8 parts x #CTAs records
        d.attn_cnt = sync.data_ptr() + 4 * 1024
        d.tile_cnt = sync.data_ptr() + 4 * 512                       # ints [512, 640): o-proj / down-proj tile counters
        d.attn_mode = int(getattr(self, "mega_attn_mode", int(os.environ.get("IVGPT_MEGA_ATTN", "0"))))
        d.a_bulk, d.mma_m64 = 1, 1
        if slot is not None:
            d.slot0, d.slot_period, d.slot_token = int(slot[0]), int(slot[1]), int(slot[2])
            d.nslots = int(slot[3].shape[1]) if slot[3] is not None else 1
            d.slot_emb = slot[3].data_ptr() if slot[3] is not None else None
        if getattr(self, "mega_profile", False):
            self.mega_prof = self.buf("mega_prof", (24,), torch.int64)
            self.mega_prof.zero_()
            d.prof = self.mega_prof.data_ptr()
        _lib.check(_lib.load().ivgpt_decode_mega(C.byref(d), torch.cuda.current_stream().cuda_stream), "decode_mega")
        return sync

    # ---- generation ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, ids: Optional[torch.Tensor], embeds: Optional[torch.Tensor], max_new_tokens: int,
                 do_sample: bool, top_k: int, temperature: float, seed: int, use_graph: bool = True,
                 use_pdl: bool = True, use_mega: Optional[bool] = None, slot_cfg=None) -> torch.Tensor:
        """Returns the token buffer [B, L + max_new_tokens] (prompt slots hold ids, or zeros for embeds).

        The decode step is captured ONCE per (batch, length, sampling mode) into a CUDA graph: position and RNG seed
        live in device memory, the token buffer is a persistent engine buffer, so later calls only replay.

        slot_cfg = (slot0, period, token, slot_emb | None): forced separator slots of the action-conditioned rollout
        (action_model.py:78-114) -- every position q >= L with (q - slot0) % period == 0 holds `token` instead of a
        sampled one and slot_emb[b, (q - slot0) // period] (fp32 [B, nslots, hidden]) is added to its embedding, so the
        whole rollout keeps ONE KV cache instead of re-prefilling the history for every frame."""
        src = ids if ids is not None else embeds
        B, L = src.shape[0], src.shape[1]
        dev = src.device
        total = L + max_new_tokens
        Lmax = (total + 7) // 8 * 8
        tokens = self.buf("tokens", (B, total), torch.int64)
        if ids is not None:
            tokens[:, :L].copy_(ids)
        else:
            tokens[:, :L].zero_()
        if max_new_tokens <= 0:
            return tokens.clone()
        V = self.w.vocab
        k = min(int(top_k), V) if (top_k is not None and top_k > 0) else V
        sample_cfg = (k, float(temperature)) if do_sample else None
        dseed = self.buf("dseed", (1,), torch.int64)
        dseed.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        dpos = self.buf("dpos", (1,), torch.int32)
        slot = None
        if slot_cfg is not None:
            s0, period, stok, semb = slot_cfg
            if semb is not None:
                assert semb.dim() == 3 and semb.shape[0] == B and semb.shape[2] == self.w.hidden
                persistent = self.buf("slot_emb", tuple(semb.shape), torch.float32)   # stable address for graph replay
                persistent.copy_(semb)
                semb = persistent
            slot = (int(s0), int(period), int(stok), semb)
        logits, _ = self.prefill(B, L, Lmax, ids.contiguous() if ids is not None else None, embeds, False)
        self._sample(logits, B, tokens, None, sample_cfg, L, dseed)       # first new token -> tokens[:, L]
        if slot is not None and L >= slot[0] and (L - slot[0]) % slot[1] == 0:
            tokens[:, L] = slot[2]
        steps = max_new_tokens - 1
        if steps == 0:
            return tokens.clone()
        dpos.fill_(L)                                                     # position of the token fed next
        if use_mega is None:
            use_mega = self.mega_supported(B, Lmax)
        if use_mega:
            if not self.mega_supported(B, Lmax):
                raise ValueError(f"decode megakernel does not support B={B}, dtype={self.dtype}, widths "
                                 f"{self.w.hidden}/{self.w.inter}")
            sync = self._decode_mega(B, Lmax, tokens, dpos, sample_cfg, dseed, steps, slot)
            out = tokens.clone()
            err = int(sync[1].item())
            if err != 0:
                raise RuntimeError({1: "decode megakernel: device-wide barrier timed out (a CTA was not co-resident?)",
                                    2: "decode megakernel: an attention ring slot never filled (bulk copy fault)",
                                    3: "decode megakernel: the activation slab never landed (bulk copy fault)",
                                    4: "decode megakernel: a weight slab never landed (bulk copy fault)",
                                    5: "decode megakernel: tensor-core completion never signalled"}.get(err, f"decode megakernel: error {err}"))
            return out
        if not use_graph:
            ops.set_pdl(use_pdl)
            try:
                for _ in range(steps):
                    self._decode_step(B, Lmax, tokens, dpos, sample_cfg, dseed, slot)
            finally:
                ops.set_pdl(False)
            return tokens.clone()
        from .. import _lib
        key = ("decode", B, Lmax, total, sample_cfg, use_pdl,
               None if slot is None else (slot[0], slot[1], slot[2], None if slot[3] is None else tuple(slot[3].shape)))
        entry = self._graphs.get(key)
        done = 0
        if entry is None:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            ops.set_pdl(use_pdl)
            try:
                with torch.cuda.stream(side):
                    self._decode_step(B, Lmax, tokens, dpos, sample_cfg, dseed, slot)   # warm-up run (is decode step 1)
                torch.cuda.current_stream(dev).wait_stream(side)
                done = 1
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    self._decode_step(B, Lmax, tokens, dpos, sample_cfg, dseed, slot)
                per_step = _lib.launch_count() - n0
                _lib.load().ivgpt_count_add(-per_step)        # capture records kernels without running them
            finally:
                ops.set_pdl(False)
            entry = (g, per_step)
            self._graphs[key] = entry
        g, per_step = entry
        for _ in range(steps - done):
            g.replay()
        _lib.load().ivgpt_count_add(per_step * (steps - done))
        return tokens.clone()
