"""Forward + backward of the Llama-style transformer on the B200 kernels (training step of reference
train_gpt.py:766-804: `outputs = model(input_ids, labels)` -> `accelerator.backward(loss)`).

Every contraction -- forward projections, dgrad, wgrad, the four attention-backward products -- is a launch of the
tcgen05 GEMM (gemm_tc.cu); both operands of that kernel are K-major, so wgrad / attention-backward operands are first
re-laid out by the batched transpose kernel.  Activations needed by the backward pass are kept per layer (B = 16 clips
of 751 tokens: ~0.6 GB per layer).  Attention dropout (config.attention_dropout, 0.1 in the reference's pre-training
scripts) is applied to the attention probabilities in training mode with a counter-based mask (ivgpt_dropout) that the
backward pass regenerates from the seed; loss/gradient parity with HF is defined at dropout 0 and, for dropout > 0,
against a torch restatement fed the SAME mask (tests/test_llama.py).
"""
from __future__ import annotations

import warnings
from typing import Dict, List, Optional

import torch

from .. import ops
from .._lib import BF16, F32
from .engine import LlamaWeights


def _r8(n: int) -> int:
    return (n + 7) // 8 * 8


def dropout_layer_seed(seed: int, layer: int) -> int:
    """Seed of the attention-dropout mask of `layer` for a forward call that drew `seed`."""
    return (int(seed) * 1000003 + layer * 7919 + 1) & 0x7FFFFFFFFFFFFFFF


class LlamaTrainEngine:
    def __init__(self, weights: LlamaWeights):
        self.w = weights
        self.dt = weights.dtype
        self.code = BF16 if self.dt == torch.bfloat16 else F32
        w = weights
        # transposed weight copies for dgrad (dX = dY . W  needs W^T as the K-major B operand)
        self.wT = []
        for lw in w.layers:
            self.wT.append({k: self._t2(lw[k]) for k in ("wqkv", "wo", "wgu", "wd")})
        self.lmT = self._t2(w.lm_head)

    # ---- small helpers ------------------------------------------------------------------------------------
    def _t2(self, x: torch.Tensor) -> torch.Tensor:
        """[R, C] -> [C, R] (row pitch padded to 8), returned as the [C, :R] view."""
        R, Cc = x.shape
        # pad columns (if any) are outside the consumer's tensor-map extent ([:, :R] is what is passed on): TMA never reads
        # them, so nothing needs zero-filling (the fills were 5 % of the step's kernel time, profiles/r02/launches_train64.txt)
        out = torch.empty(Cc, _r8(R), dtype=x.dtype, device=x.device)
        ops.transpose_raw(x, 0, out, 1, R, Cc, x.stride(0), out.stride(0), 0, 0)
        return out[:, :R]

    def _cast(self, x32: torch.Tensor) -> torch.Tensor:
        return ops.convert(x32, self.dt)

    def _wgrad(self, dY: torch.Tensor, X: torch.Tensor) -> torch.Tensor:
        """dW [N, K] = dY[M, N]^T . X[M, K]   (fp32 output)"""
        return ops.gemm(self._t2(dY), self._t2(X), out_dtype=torch.float32)

    # ---- forward + backward --------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_backward(self, ids: Optional[torch.Tensor], labels: torch.Tensor, grad_scale: float = 1.0, on_grads=None,
                         embeds: Optional[torch.Tensor] = None, attn_dropout: float = 0.0, seed: int = 0):
        """Returns (loss 0-d fp32 tensor, grads dict keyed by HF parameter name -> fp32 tensor).

        embeds [B, L, hidden] instead of ids: the action-conditioned training entry (HeadModelWithAction.forward,
        reference action_model.py:171-186); the gradient w.r.t. the embeddings is returned as grads["inputs_embeds"]
        (fp32 [B, L, hidden]) and no embedding-table gradient is produced here (autograd routes it through the caller's
        embedding lookup).  attn_dropout > 0: dropout on the attention probabilities (HF LlamaAttention in training mode),
        mask regenerated from (seed, layer) in the backward pass.

        on_grads(grads, names): called as soon as the gradients `names` are final, in backward order ([lm_head, final norm],
        layer L-1 ... layer 0, [embedding]) -- the hook of the data-parallel exchange (grad_reduce.BucketedGradReducer
        launches that bucket's all-reduce while the next layer's backward is computed).  The hook may replace entries."""
        w, dt, code = self.w, self.dt, self.code
        src = ids if ids is not None else embeds
        B, L = src.shape[0], src.shape[1]
        M, h, H, I, V = B * L, w.hidden, w.heads, w.inter, w.vocab
        Lp = _r8(L)
        dev = src.device
        labels = labels.contiguous()
        if ids is not None:
            ids = ids.contiguous()
            x = torch.empty(M, h, dtype=torch.float32, device=dev)
            ops.embed(ids, ids.stride(0), L, None, w.embed, x, M)
        else:
            assert embeds.shape == (B, L, h), (embeds.shape, (B, L, h))
            x = embeds.detach().to(torch.float32).reshape(M, h).clone()
        p_drop = float(attn_dropout)
        layer_seed = lambda li: dropout_layer_seed(seed, li)
        saved: List[Dict[str, torch.Tensor]] = []
        kc = torch.zeros(B, H, Lp, 64, dtype=dt, device=dev)
        vc = torch.zeros(B, H, 64, Lp, dtype=dt, device=dev)
        for li, lw in enumerate(w.layers):
            s: Dict[str, torch.Tensor] = {"xin": x.clone()}
            xn1 = torch.empty(M, h, dtype=dt, device=dev)
            ops.rmsnorm(x, lw["n1"], xn1, M, w.eps)
            qkv = ops.gemm(xn1, lw["wqkv"])
            q = torch.empty(B, H, L, 64, dtype=dt, device=dev)
            k = torch.zeros(B, H, Lp, 64, dtype=dt, device=dev)
            vt = torch.zeros(B, H, 64, Lp, dtype=dt, device=dev)
            ops.rope_kv(qkv, q, k, vt, B, L, H, Lp, 0, None, w.cos, w.sin)
            sc = torch.empty(B * H, L, Lp, dtype=torch.float32, device=dev)
            ops.gemm_raw(ops.gemm_desc(
                dtype=code, a=q.data_ptr(), lda=64, a_bstride=L * 64, a_rows=L, a_cols=64, a_batches=B * H,
                b=k.data_ptr(), ldb=64, b_bstride=Lp * 64, b_rows=L, b_cols=64, b_batches=B * H,
                M=L, N=L, K=64, batch=B * H, heads=1, a_bsel=2, b_bsel=2, o_bsel=2, causal_skip=1,
                out=sc.data_ptr(), ldo=Lp, out_bstride=L * Lp, out_dtype=F32, alpha=0.125))
            P = torch.empty(B * H, L, Lp, dtype=dt, device=dev)
            ops.softmax(sc, P, B * H * L, L, L, Lp, Lp, True, 0)
            Pd = ops.dropout(P, p_drop, layer_seed(li)) if p_drop > 0.0 else P      # P' = dropout(P); P itself is kept
            ao = torch.empty(M, h, dtype=dt, device=dev)
            ops.gemm_raw(ops.gemm_desc(
                dtype=code, a=Pd.data_ptr(), lda=Lp, a_bstride=L * Lp, a_rows=L, a_cols=L, a_batches=B * H,
                b=vt.data_ptr(), ldb=Lp, b_bstride=64 * Lp, b_rows=64, b_cols=L, b_batches=B * H,
                M=L, N=64, K=L, batch=B * H, heads=H, a_bsel=2, b_bsel=2, o_bsel=1, o_nhead=64,
                out=ao.data_ptr(), ldo=h, out_bstride=L * h, out_dtype=code))
            ops.gemm(ao, lw["wo"], residual=x, out=x)
            s["xmid"] = x.clone()
            xn2 = torch.empty(M, h, dtype=dt, device=dev)
            ops.rmsnorm(x, lw["n2"], xn2, M, w.eps)
            gu = ops.gemm(xn2, lw["wgu"])
            act = ops.swiglu(gu)
            ops.gemm(act, lw["wd"], residual=x, out=x)
            del Pd
            s.update(xn1=xn1, qkv=qkv, q=q, k=k, P=P, ao=ao, xn2=xn2, gu=gu, act=act)
            saved.append(s)
        xnf = torch.empty(M, h, dtype=dt, device=dev)
        ops.rmsnorm(x, w.norm, xnf, M, w.eps)
        Vp = _r8(V)
        logits = torch.empty(B, L, Vp, dtype=torch.float32, device=dev)
        ops.gemm(xnf, w.lm_head, out=logits.view(M, Vp)[:, :V])
        loss, _, count = ops.ce_loss(logits, Vp, B, L, V, labels)

        # ---------------- backward ----------------
        grads: Dict[str, torch.Tensor] = {}
        dlogits = torch.empty(B, L, Vp, dtype=dt, device=dev)
        ops.ce_bwd(logits, Vp, B, L, V, labels, count, float(grad_scale), dlogits)
        dl2 = dlogits.view(M, Vp)[:, :V]
        grads["lm_head.weight"] = self._wgrad(dl2, xnf)
        d_xnf = ops.gemm(dl2, self.lmT)
        del logits, dlogits
        dx = torch.zeros(M, h, dtype=torch.float32, device=dev)
        grads["model.norm.weight"] = ops.rmsnorm_bwd(x, w.norm, d_xnf, dx, w.eps)
        if on_grads is not None:
            on_grads(grads, ["lm_head.weight", "model.norm.weight"])
        for li in reversed(range(w.layers_n)):
            lw, wt, s = w.layers[li], self.wT[li], saved[li]
            pre = f"model.layers.{li}."
            # ---- MLP ----
            dxT = self._cast(dx)
            d_act = ops.gemm(dxT, wt["wd"])
            grads[pre + "mlp.down_proj.weight"] = self._wgrad(dxT, s["act"])
            d_gu = ops.swiglu(s["gu"], d_act)
            d_xn2 = ops.gemm(d_gu, wt["wgu"])
            g_gu = self._wgrad(d_gu, s["xn2"])
            grads[pre + "mlp.gate_proj.weight"] = g_gu[0::2]
            grads[pre + "mlp.up_proj.weight"] = g_gu[1::2]
            grads[pre + "post_attention_layernorm.weight"] = ops.rmsnorm_bwd(s["xmid"], lw["n2"], d_xn2, dx, w.eps)
            # ---- attention ----
            dxT = self._cast(dx)
            d_ao = ops.gemm(dxT, wt["wo"])
            grads[pre + "self_attn.o_proj.weight"] = self._wgrad(dxT, s["ao"])
            qkv, P = s["qkv"], s["P"]
            dP = torch.empty(B * H, L, Lp, dtype=torch.float32, device=dev)
            ops.gemm_raw(ops.gemm_desc(      # dP[b,h] = dO_h . V_h^T      (V_h read in place from the fused qkv buffer)
                dtype=code, a=d_ao.data_ptr(), lda=h, a_bstride=L * h, a_rows=L, a_cols=h, a_batches=B,
                b=qkv.data_ptr(), ldb=3 * h, b_bstride=L * 3 * h, b_rows=L, b_cols=3 * h, b_batches=B,
                M=L, N=L, K=64, batch=B * H, heads=H, a_bsel=1, a_bdiv=1, b_bsel=1, b_bdiv=1, o_bsel=2,
                a_khead=64, b_kbase=2 * h, b_khead=64, causal_skip=1,
                out=dP.data_ptr(), ldo=Lp, out_bstride=L * Lp, out_dtype=F32))
            if p_drop > 0.0:
                ops.dropout(dP, p_drop, layer_seed(li), out=dP)          # dP = dP' * mask / (1 - p), same mask as the forward
            dS = torch.empty(B * H, L, Lp, dtype=dt, device=dev)
            ops.softmax_bwd(P, dP, dS, B * H * L, L, L, Lp, True, 0.125)
            del dP
            Pt = torch.empty(B * H, L, Lp, dtype=dt, device=dev)       # read through maps of extent L: the pad column is never touched
            Pd = ops.dropout(P, p_drop, layer_seed(li)) if p_drop > 0.0 else P      # dV = P'^T . dO
            ops.transpose_raw(Pd, 0, Pt, B * H, L, L, Lp, Lp, L * Lp, L * Lp)
            del Pd
            d_aoT = torch.empty(B, h, Lp, dtype=dt, device=dev)       # read through maps of extent L: the pad column is never touched
            ops.transpose_raw(d_ao, 0, d_aoT, B, L, h, h, Lp, L * h, h * Lp)
            dV = torch.empty(B, H, L, 64, dtype=torch.float32, device=dev)
            ops.gemm_raw(ops.gemm_desc(      # dV[b,h] = P^T . dO_h
                dtype=code, a=Pt.data_ptr(), lda=Lp, a_bstride=L * Lp, a_rows=L, a_cols=L, a_batches=B * H,
                b=d_aoT.data_ptr(), ldb=Lp, b_bstride=h * Lp, b_rows=h, b_cols=L, b_batches=B,
                M=L, N=64, K=L, batch=B * H, heads=H, a_bsel=2, b_bsel=1, b_bdiv=1, o_bsel=2, b_nhead=64,
                out=dV.data_ptr(), ldo=64, out_bstride=L * 64, out_dtype=F32))
            del Pt
            kT = torch.empty(B * H, 64, Lp, dtype=dt, device=dev)       # read through maps of extent L: the pad column is never touched
            ops.transpose_raw(s["k"], 0, kT, B * H, L, 64, 64, Lp, Lp * 64, 64 * Lp)
            dQ = torch.empty(B, H, L, 64, dtype=torch.float32, device=dev)
            ops.gemm_raw(ops.gemm_desc(      # dQ'[b,h] = dS . K'
                dtype=code, a=dS.data_ptr(), lda=Lp, a_bstride=L * Lp, a_rows=L, a_cols=L, a_batches=B * H,
                b=kT.data_ptr(), ldb=Lp, b_bstride=64 * Lp, b_rows=64, b_cols=L, b_batches=B * H,
                M=L, N=64, K=L, batch=B * H, heads=1, a_bsel=2, b_bsel=2, o_bsel=2,
                out=dQ.data_ptr(), ldo=64, out_bstride=L * 64, out_dtype=F32))
            dSt = torch.empty(B * H, L, Lp, dtype=dt, device=dev)       # read through maps of extent L: the pad column is never touched
            ops.transpose_raw(dS, 0, dSt, B * H, L, L, Lp, Lp, L * Lp, L * Lp)
            qT = torch.empty(B * H, 64, Lp, dtype=dt, device=dev)       # read through maps of extent L: the pad column is never touched
            ops.transpose_raw(s["q"], 0, qT, B * H, L, 64, 64, Lp, L * 64, 64 * Lp)
            dK = torch.empty(B, H, L, 64, dtype=torch.float32, device=dev)
            ops.gemm_raw(ops.gemm_desc(      # dK'[b,h] = dS^T . Q'
                dtype=code, a=dSt.data_ptr(), lda=Lp, a_bstride=L * Lp, a_rows=L, a_cols=L, a_batches=B * H,
                b=qT.data_ptr(), ldb=Lp, b_bstride=64 * Lp, b_rows=64, b_cols=L, b_batches=B * H,
                M=L, N=64, K=L, batch=B * H, heads=1, a_bsel=2, b_bsel=2, o_bsel=2,
                out=dK.data_ptr(), ldo=64, out_bstride=L * 64, out_dtype=F32))
            del dS, dSt
            d_qkv = torch.empty(M, 3 * h, dtype=dt, device=dev)
            ops.rope_bwd(dQ, dK, dV, d_qkv, B, L, H, w.cos, w.sin)
            d_xn1 = ops.gemm(d_qkv, wt["wqkv"])
            g_qkv = self._wgrad(d_qkv, s["xn1"])
            grads[pre + "self_attn.q_proj.weight"] = g_qkv[0:h]
            grads[pre + "self_attn.k_proj.weight"] = g_qkv[h:2 * h]
            grads[pre + "self_attn.v_proj.weight"] = g_qkv[2 * h:3 * h]
            grads[pre + "input_layernorm.weight"] = ops.rmsnorm_bwd(s["xin"], lw["n1"], d_xn1, dx, w.eps)
            saved[li] = None
            if on_grads is not None:
                on_grads(grads, [pre + n for n in (
                    "mlp.down_proj.weight", "mlp.gate_proj.weight", "mlp.up_proj.weight", "post_attention_layernorm.weight",
                    "self_attn.o_proj.weight", "self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight",
                    "input_layernorm.weight")])
        if ids is not None:
            dE = torch.zeros(V, h, dtype=torch.float32, device=dev)
            ops.embed_bwd(ids, dx, dE)
            grads["model.embed_tokens.weight"] = dE
            if on_grads is not None:
                on_grads(grads, ["model.embed_tokens.weight"])
        else:
            grads["inputs_embeds"] = dx.view(B, L, h)
        return loss, grads
