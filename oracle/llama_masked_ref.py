"""ORACLE (test infrastructure, NOT product code) -- Llama forward with an EXPLICIT attention-dropout mask.

HF's LlamaAttention applies `nn.functional.dropout(attn_weights, p=attention_dropout, training=self.training)` to the
softmax output (transformers models/llama/modeling_llama.py, eager_attention_forward; the reference trains with
`--attention_dropout 0.1`, scripts/pretrain/oxe-64-act-free.sh:31).  torch's dropout mask cannot be reproduced by another
generator, so parity at dropout > 0 is defined as: GIVEN THE SAME MASK, loss and gradients agree.  This module restates
the HF Llama forward in plain torch (fp32, autograd) over an HF state dict, taking the per-layer multiplicative mask
(keep / (1 - p), or zeros) as an input; with all-ones masks it must reproduce the unmodified HF model exactly
(checked in tests/test_llama.py on CPU), which pins the restatement to the genuine class.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


def _rms(x, w, eps):
    v = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(v + eps))


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


def masked_llama_loss(sd: Dict[str, torch.Tensor], cfg, ids: Optional[torch.Tensor], labels: torch.Tensor,
                      masks: Optional[List[torch.Tensor]] = None, embeds: Optional[torch.Tensor] = None):
    """sd: HF LlamaForCausalLM state dict (tensors may require grad); masks[l]: [B, heads, L, L] multiplier applied to
    the attention probabilities of layer l (None = no dropout).  Returns (loss, logits)."""
    H = cfg.num_attention_heads
    h = cfg.hidden_size
    hd = h // H
    eps = cfg.rms_norm_eps
    theta = float(getattr(cfg, "rope_theta", None) or (getattr(cfg, "rope_parameters", None) or {}).get("rope_theta", 10000.0))
    x = F.embedding(ids, sd["model.embed_tokens.weight"]) if embeds is None else embeds
    B, L, _ = x.shape
    inv = 1.0 / (theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.arange(L, dtype=torch.float32)[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], dim=-1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]
    causal = torch.full((L, L), float("-inf")).triu(1)
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        r = x
        y = _rms(x, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(y, sd[p + "self_attn.q_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        k = F.linear(y, sd[p + "self_attn.k_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        v = F.linear(y, sd[p + "self_attn.v_proj.weight"]).view(B, L, H, hd).transpose(1, 2)
        q = q * cos + _rotate_half(q) * sin
        k = k * cos + _rotate_half(k) * sin
        s = q @ k.transpose(-1, -2) / (hd ** 0.5) + causal
        pr = torch.softmax(s, dim=-1, dtype=torch.float32)
        if masks is not None and masks[l] is not None:
            pr = pr * masks[l]
        a = (pr @ v).transpose(1, 2).reshape(B, L, h)
        x = r + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        r = x
        y = _rms(x, sd[p + "post_attention_layernorm.weight"], eps)
        g = F.linear(y, sd[p + "mlp.gate_proj.weight"])
        u = F.linear(y, sd[p + "mlp.up_proj.weight"])
        x = r + F.linear(F.silu(g) * u, sd[p + "mlp.down_proj.weight"])
    x = _rms(x, sd["model.norm.weight"], eps)
    logits = F.linear(x, sd["lm_head.weight"])
    loss = F.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]).float(), labels[:, 1:].reshape(-1), ignore_index=-100)
    return loss, logits
