"""ORACLE (test infrastructure, NOT product code) -- the transformer half of the path.

Unlike the tokenizer, the reference's transformer arithmetic can be executed here: it is transformers'
`LlamaForCausalLM` (requirements.txt:4 pins 4.38.2; this image has 5.5.0 -- version skew stated in DESIGN.md),
reached by reference inference/predict.py:111 / train_gpt.py:597.  This module builds the UNMODIFIED HF class in
fp32 with eager attention on CPU, with seeded weights, and drives it the way predict.py:54-69 does.
"""
from __future__ import annotations

import json
import os

import torch


def config_path(name: str) -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", name + ".json")


def build_hf_llama(cfg: dict | str, seed: int = 4321, init_scale: float = 1.0):
    """Genuine transformers.LlamaForCausalLM (never the B200 subclass), fp32, eval, seeded N(0, 0.02*scale)."""
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaForCausalLM
    if isinstance(cfg, str):
        with open(cfg) as fh:
            cfg = json.load(fh)
    cfg = dict(cfg)
    cfg.pop("torch_dtype", None), cfg.pop("architectures", None)
    config = LlamaConfig(**cfg)
    config._attn_implementation = "eager"
    g = torch.Generator().manual_seed(seed)
    model = LlamaForCausalLM(config).to(torch.float32).eval()
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * (0.02 * init_scale))
            else:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
    return model


TINY_LLAMA = dict(model_type="llama", hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                  num_attention_heads=2, num_key_value_heads=2, hidden_act="silu", max_position_embeddings=1024,
                  rms_norm_eps=1e-6, tie_word_embeddings=False, vocab_size=1026, bos_token_id=50256,
                  eos_token_id=50256, use_cache=True)


@torch.no_grad()
def greedy_generate(model, ids: torch.Tensor, max_new_tokens: int) -> torch.Tensor:
    """predict.py:64-69 with do_sample=False."""
    return model.generate(ids, do_sample=False, max_new_tokens=max_new_tokens, pad_token_id=50256)
