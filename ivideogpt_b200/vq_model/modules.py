"""Parameter containers of the compressive tokenizer.

These modules exist to own nn.Parameters under exactly the names the reference's checkpoints use
(``diffusion_pytorch_model.safetensors`` written by diffusers ModelMixin for the module tree built in
reference ivideogpt/vq_model/vae.py:84-137,234-294 and conditional_vae.py:88-106,166-184), so that
``load_state_dict(strict=True)`` works on released weights.  None of them has a forward(): all arithmetic
is issued by ``plan.py`` as B200 kernels on packed copies of these parameters.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn


class _NoForward(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - guard against accidental eager use
        raise RuntimeError(f"{type(self).__name__} is a parameter container; the B200 tokenizer runs through "
                           "ivideogpt_b200.vq_model.plan, not through nn.Module.forward")


class ResnetParams(_NoForward):
    """norm1 -> silu -> conv1 -> norm2 -> silu -> conv2 (+ 1x1 conv_shortcut when channels change)."""

    def __init__(self, cin: int, cout: int, groups: int):
        super().__init__()
        self.cin, self.cout = cin, cout
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None


class _Resampler(_NoForward):
    def __init__(self, ch: int, stride: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=stride, padding=0 if stride == 2 else 1)


class StageParams(_NoForward):
    """One encoder (down) or decoder (up) stage: a run of resnets + optional 3x3 resampling conv."""

    def __init__(self, cin: int, cout: int, n_res: int, resample: str, groups: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetParams(cin if j == 0 else cout, cout, groups) for j in range(n_res)])
        self.downsamplers = None
        self.upsamplers = None
        if resample == "down":
            self.downsamplers = nn.ModuleList([_Resampler(cout, 2)])
        elif resample == "up":
            self.upsamplers = nn.ModuleList([_Resampler(cout, 1)])


class MidAttnParams(_NoForward):
    """Single-head spatial self-attention of the mid block (diffusers Attention key layout)."""

    def __init__(self, ch: int, groups: int):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Identity()])


class MidParams(_NoForward):
    def __init__(self, ch: int, groups: int, attention: bool):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetParams(ch, ch, groups), ResnetParams(ch, ch, groups)])
        self.attentions = nn.ModuleList([MidAttnParams(ch, groups)]) if attention else nn.ModuleList([])
        self.has_attention = attention


class _MHAParams(_NoForward):
    """Key layout of nn.MultiheadAttention: in_proj_weight [3C,C], in_proj_bias, out_proj.{weight,bias}."""

    def __init__(self, ch: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * ch, ch))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * ch))
        self.out_proj = nn.Linear(ch, ch)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class CrossAttnParams(_NoForward):
    """conditional_vae.py:10-36: 4-head attention from frame tokens to context-frame tokens."""

    def __init__(self, ch: int, res: int, kv_frames: int, groups: int = 32, heads: int = 4):
        super().__init__()
        self.ch, self.res, self.heads = ch, res, heads
        self.dropout = 0.1   # attention-probability and residual dropout of the reference block (train mode only; :17,24-25)
        self.kv_frames = kv_frames
        self.att = _MHAParams(ch)
        self.kv_norm = nn.GroupNorm(groups, ch)  # eps 1e-5 (torch default), unlike the resnet norms
        self.q_norm = nn.GroupNorm(groups, ch)
        self.kv_pos_emb = nn.Parameter(torch.zeros(kv_frames * res * res, ch))
        self.q_pos_emb = nn.Parameter(torch.zeros(res * res, ch))

    def set_kv_frames(self, kv_frames: int):
        """Keep the positional embeddings of the LAST kv_frames context frames (conditional_vae.py:34-36)."""
        keep = kv_frames * self.kv_pos_emb.shape[0] // self.kv_frames
        self.kv_pos_emb.data = self.kv_pos_emb.data[-keep:]
        self.kv_frames = kv_frames


class EncoderParams(_NoForward):
    """vae.py:72-137 (+ cross-attention blocks of conditional_vae.py:88-106 when `cross` is set)."""

    def __init__(self, in_ch: int, out_ch: int, chans: Sequence[int], layers: int, groups: int, mid_attention: bool,
                 cross: bool = False, max_att: int = 0, init_res: int = 0, ctx: int = 1):
        super().__init__()
        self.chans = tuple(chans)
        self.conv_in = nn.Conv2d(in_ch, chans[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = chans[0]
        for i, co in enumerate(chans):
            last = i == len(chans) - 1
            self.down_blocks.append(StageParams(c, co, layers, "none" if last else "down", groups))
            c = co
        self.mid_block = MidParams(chans[-1], groups, mid_attention)
        self.conv_norm_out = nn.GroupNorm(groups, chans[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[-1], out_ch, 3, padding=1)
        self.max_att_resolution = max_att
        if cross:
            self.cross_att_blocks = nn.ModuleList()
            res = init_res
            for i, co in enumerate(chans):
                if i != len(chans) - 1:
                    res //= 2
                if res <= max_att:
                    self.cross_att_blocks.append(CrossAttnParams(co, res, ctx, groups))

    def set_context_length(self, n: int):
        for blk in self.cross_att_blocks:
            blk.set_kv_frames(n)


class DecoderParams(_NoForward):
    """vae.py:221-294 (+ cross-attention blocks of conditional_vae.py:166-184 when `cross` is set)."""

    def __init__(self, in_ch: int, out_ch: int, chans: Sequence[int], layers: int, groups: int, mid_attention: bool,
                 cross: bool = False, max_att: int = 0, init_res: int = 16, ctx: int = 1):
        super().__init__()
        self.chans = tuple(chans)
        rev = list(reversed(chans))
        self.conv_in = nn.Conv2d(in_ch, chans[-1], 3, padding=1)
        self.mid_block = MidParams(chans[-1], groups, mid_attention)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            last = i == len(chans) - 1
            self.up_blocks.append(StageParams(c, co, layers + 1, "none" if last else "up", groups))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, chans[0], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[0], out_ch, 3, padding=1)
        self.max_att_resolution = max_att
        if cross:
            res = init_res
            self.cross_att_blocks = nn.ModuleList([CrossAttnParams(rev[0], res, ctx, groups)])
            for i, co in enumerate(rev):
                if i != len(chans) - 1:
                    res *= 2
                if res <= max_att:
                    self.cross_att_blocks.append(CrossAttnParams(co, res, ctx, groups))

    def set_context_length(self, n: int):
        for blk in self.cross_att_blocks:
            blk.set_kv_frames(n)


class Codebook(_NoForward):
    """Key layout of diffusers VectorQuantizer: `embedding.weight` [n_e, dim], init U(-1/n_e, 1/n_e)."""

    def __init__(self, n_e: int, dim: int):
        super().__init__()
        self.n_e, self.vq_embed_dim = n_e, dim
        self.embedding = nn.Embedding(n_e, dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)
