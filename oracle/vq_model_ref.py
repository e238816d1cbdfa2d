"""ORACLE (test infrastructure, NOT product code) -- CPU fp32 restatement of the iVideoGPT
compressive tokenizer.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package ``ivideogpt_b200``
never does.

PARITY, what is pinned and what is not.  The reference ships no golden vectors for this path and the block-level
arithmetic lives in ``diffusers==0.27.0`` (requirements.txt:5), which is installed neither here nor on the GPU box.
  * PINNED -- the reference's own code: tests/golden/make_golden_tokenizer_ref.py imports and RUNS the reference's
    vae.py / conditional_vae.py / compressive_vq_model.py unmodified (on top of oracle/diffusers_stub, stand-ins for the
    diffusers symbols they import) for the tiny config and the full ctx_vae64 / ctx_vae256 configs; this restatement
    reproduces those runs exactly (state-dict keys, tokens, labels, reconstructed pixels: max abs diff 0.0), see
    tests/golden/tokenizer_refglue.npz and tests/test_tokenizer.py.
  * STILL RESTATED ("parity unpinned" at block level) -- the diffusers building blocks themselves (ResnetBlock2D,
    Down/UpDecoderBlock2D, UNetMidBlock2D, Attention, VectorQuantizer): the classes below restate the library's published
    algorithm and are the same classes the stub hands to the reference's code; their structure is pinned by
    exact state-dict key/shape agreement with the reference's module tree and the README's parameter counts
    (114.2 M / 310.5 M).

What is restated, with the reference lines it follows:
  * ResnetBlock2D / DownEncoderBlock2D / Downsample2D / UpDecoderBlock2D / Upsample2D /
    UNetMidBlock2D / Attention / VectorQuantizer   -- diffusers 0.27.0, used at
    ivideogpt/vq_model/vae.py:104-130,250-285 and compressive_vq_model.py:102-123
  * Encoder / Decoder                               -- ivideogpt/vq_model/vae.py:47-195,198-371
  * CrossAttentionBlock / Conditional{En,De}coder   -- ivideogpt/vq_model/conditional_vae.py:10-212
  * CompressiveVQModel.tokenize / detokenize        -- ivideogpt/vq_model/compressive_vq_model.py:165-277
"""
from __future__ import annotations

import json
import math
import os
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# diffusers 0.27.0 blocks (restated)
# --------------------------------------------------------------------------------------
class RefResnetBlock(nn.Module):
    """diffusers ResnetBlock2D(temb_channels=None, groups=32, eps=1e-6, silu, scale 1)."""

    def __init__(self, cin: int, cout: int, groups: int = 32, eps: float = 1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class _ConvHolder(nn.Module):
    """Gives the `downsamplers.0.conv` / `upsamplers.0.conv` key path."""

    def __init__(self, conv: nn.Conv2d):
        super().__init__()
        self.conv = conv


class RefDownBlock(nn.Module):
    """diffusers DownEncoderBlock2D(num_layers, add_downsample, downsample_padding=0)."""

    def __init__(self, cin, cout, num_layers, add_downsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList(
            [RefResnetBlock(cin if j == 0 else cout, cout, groups) for j in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, 2, 0))])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.downsamplers is not None:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0.0)
            x = self.downsamplers[0].conv(x)
        return x


class RefUpBlock(nn.Module):
    """diffusers UpDecoderBlock2D(num_layers, add_upsample): resnets, nearest-2x, 3x3 conv."""

    def __init__(self, cin, cout, num_layers, add_upsample, groups):
        super().__init__()
        self.resnets = nn.ModuleList(
            [RefResnetBlock(cin if j == 0 else cout, cout, groups) for j in range(num_layers)])
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([_ConvHolder(nn.Conv2d(cout, cout, 3, 1, 1))])

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.upsamplers is not None:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = self.upsamplers[0].conv(x)
        return x


class RefMidAttention(nn.Module):
    """diffusers Attention as built by UNetMidBlock2D: 1 head of dim C, GroupNorm(32,C,1e-6),
    biased q/k/v/out projections, residual connection, scale 1/sqrt(C)."""

    def __init__(self, c, groups, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=eps, affine=True)
        self.to_q = nn.Linear(c, c)
        self.to_k = nn.Linear(c, c)
        self.to_v = nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        a = torch.softmax((q @ k.transpose(1, 2)) * (1.0 / math.sqrt(c)), dim=-1) @ v
        a = self.to_out[0](a)
        return a.transpose(1, 2).reshape(b, c, h, w) + x


class RefMidBlock(nn.Module):
    """diffusers UNetMidBlock2D(num_layers=1): resnet, [attention], resnet."""

    def __init__(self, c, groups, add_attention):
        super().__init__()
        self.resnets = nn.ModuleList([RefResnetBlock(c, c, groups), RefResnetBlock(c, c, groups)])
        self.attentions = nn.ModuleList([RefMidAttention(c, groups) if add_attention else None])
        self.add_attention = add_attention

    def forward(self, x):
        x = self.resnets[0](x)
        if self.attentions[0] is not None:
            x = self.attentions[0](x)
        return self.resnets[1](x)


class RefVectorQuantizer(nn.Module):
    """diffusers VectorQuantizer(beta=1.0, legacy=False, sane_index_shape=False)."""

    def __init__(self, n_e, dim):
        super().__init__()
        self.n_e, self.dim = n_e, dim
        self.embedding = nn.Embedding(n_e, dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def forward(self, z_nchw, beta: float = 1.0, idx=None):
        """diffusers VectorQuantizer.forward(legacy=False): (straight-through z_q NCHW, loss, indices).  Differentiable.
        idx (test aid, not in diffusers): use these indices instead of the argmin -- lets a test hold the discrete choice fixed
        when comparing an implementation whose latents differ by rounding (an index may flip on a near-tie)."""
        z = z_nchw.permute(0, 2, 3, 1).contiguous()
        zf = z.view(-1, self.dim)
        if idx is None:
            idx = torch.argmin(torch.cdist(zf, self.embedding.weight), dim=1)
        z_q = self.embedding(idx).view(z.shape)
        loss = beta * torch.mean((z_q.detach() - z) ** 2) + torch.mean((z_q - z.detach()) ** 2)
        z_q = z + (z_q - z).detach()
        return z_q.permute(0, 3, 1, 2).contiguous(), loss, idx

    def indices(self, z_nchw):
        zf = z_nchw.permute(0, 2, 3, 1).contiguous().view(-1, self.dim)
        return torch.argmin(torch.cdist(zf, self.embedding.weight), dim=1)


# --------------------------------------------------------------------------------------
# reference modules (vae.py / conditional_vae.py) restated
# --------------------------------------------------------------------------------------
class RefEncoder(nn.Module):
    """vae.py:47-195 (double_z=False)."""

    def __init__(self, in_ch, out_ch, chans: Sequence[int], layers_per_block, groups, mid_attn):
        super().__init__()
        self.conv_in = nn.Conv2d(in_ch, chans[0], 3, 1, 1)
        self.down_blocks = nn.ModuleList()
        c = chans[0]
        for i, co in enumerate(chans):
            self.down_blocks.append(RefDownBlock(c, co, layers_per_block, i != len(chans) - 1, groups))
            c = co
        self.mid_block = RefMidBlock(chans[-1], groups, mid_attn)
        self.conv_norm_out = nn.GroupNorm(groups, chans[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[-1], out_ch, 3, padding=1)

    def forward(self, x, return_features=False):
        feats = []
        x = self.conv_in(x)
        feats.append(x)
        for blk in self.down_blocks:
            x = blk(x)
            feats.append(x)
        x = self.mid_block(x)
        feats.append(x)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x, feats) if return_features else x


class RefDecoder(nn.Module):
    """vae.py:198-371 (norm_type='group')."""

    def __init__(self, in_ch, out_ch, chans: Sequence[int], layers_per_block, groups, mid_attn):
        super().__init__()
        rev = list(reversed(chans))
        self.conv_in = nn.Conv2d(in_ch, chans[-1], 3, 1, 1)
        self.mid_block = RefMidBlock(chans[-1], groups, mid_attn)
        self.up_blocks = nn.ModuleList()
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(RefUpBlock(c, co, layers_per_block + 1, i != len(chans) - 1, groups))
            c = co
        self.conv_norm_out = nn.GroupNorm(groups, chans[0], eps=1e-6)
        self.conv_out = nn.Conv2d(chans[0], out_ch, 3, padding=1)

    def forward(self, x, return_features=False):
        feats = []
        x = self.conv_in(x)
        feats.append(x)
        x = self.mid_block(x)
        feats.append(x)
        for blk in self.up_blocks:
            x = blk(x)
            feats.append(x)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x, feats) if return_features else x


class RefCrossAttention(nn.Module):
    """conditional_vae.py:10-55 in eval mode (dropouts are identity)."""

    def __init__(self, c, res, kv_frames, groups=32, heads=4):
        super().__init__()
        self.att = nn.MultiheadAttention(c, heads, dropout=0.1, batch_first=True)
        self.kv_norm = nn.GroupNorm(groups, c)
        self.q_norm = nn.GroupNorm(groups, c)
        self.kv_frames = kv_frames
        self.kv_pos_emb = nn.Parameter(torch.zeros(kv_frames * res * res, c))
        self.q_pos_emb = nn.Parameter(torch.zeros(res * res, c))

    def forward(self, z, addin):
        if self.kv_frames > 1:  # [B,t,C,H,W] -> [B,C,tH,W]
            addin = addin.permute(0, 2, 1, 3, 4).reshape(addin.shape[0], addin.shape[2], -1, addin.shape[3])
        kv = self.kv_norm(addin).permute(0, 2, 3, 1).reshape(addin.shape[0], -1, addin.shape[1]) + self.kv_pos_emb
        q = self.q_norm(z).permute(0, 2, 3, 1).reshape(z.shape[0], -1, z.shape[1]) + self.q_pos_emb
        out, _ = self.att(q, kv, kv)
        out = out.permute(0, 2, 1).reshape(z.shape)
        return F.silu(z + out)


class RefCondEncoder(RefEncoder):
    """conditional_vae.py:58-132."""

    def __init__(self, in_ch, out_ch, chans, layers_per_block, groups, max_att, init_res, ctx):
        super().__init__(in_ch, out_ch, chans, layers_per_block, groups, True)
        self.max_att = max_att
        self.cross_att_blocks = nn.ModuleList()
        res = init_res
        for i, co in enumerate(chans):
            if i != len(chans) - 1:
                res //= 2
            if res <= max_att:
                self.cross_att_blocks.append(RefCrossAttention(co, res, ctx, groups))

    def forward(self, x, cond):
        x = self.conv_in(x)
        k = 0
        for i, blk in enumerate(self.down_blocks):
            x = blk(x)
            if x.shape[-1] <= self.max_att:
                x = self.cross_att_blocks[k](x, cond[i + 1])
                k += 1
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class RefCondDecoder(RefDecoder):
    """conditional_vae.py:135-212."""

    def __init__(self, in_ch, out_ch, chans, layers_per_block, groups, max_att, init_res, ctx):
        super().__init__(in_ch, out_ch, chans, layers_per_block, groups, True)
        self.max_att = max_att
        rev = list(reversed(chans))
        res = init_res
        self.cross_att_blocks = nn.ModuleList([RefCrossAttention(rev[0], res, ctx, groups)])
        for i, co in enumerate(rev):
            if i != len(chans) - 1:
                res *= 2
            if res <= max_att:
                self.cross_att_blocks.append(RefCrossAttention(co, res, ctx, groups))

    def forward(self, x, cond):
        x = self.conv_in(x)
        x = self.mid_block(x)
        x = self.cross_att_blocks[0](x, cond[1])
        for i, blk in enumerate(self.up_blocks):
            x = blk(x)
            if x.shape[-1] <= self.max_att:
                x = self.cross_att_blocks[i + 1](x, cond[i + 2])
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class RefCompressiveVQModel(nn.Module):
    """compressive_vq_model.py:33-277 (inference methods only)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(64,), layers_per_block=1,
                 latent_channels=3, num_vq_embeddings=256, norm_num_groups=32, vq_embed_dim=None,
                 mid_block_add_attention=True, num_dyn_embeddings=256, context_length=1,
                 max_att_resolution=32, resolution=256, patch_size=4, **_unused):
        super().__init__()
        ch = tuple(block_out_channels)
        self.context_length = context_length
        self.num_vq_embeddings = num_vq_embeddings
        self.num_dyn_embeddings = num_dyn_embeddings
        self.patch_size = patch_size
        self.latent_channels = latent_channels
        vq_embed_dim = vq_embed_dim if vq_embed_dim is not None else latent_channels
        self.vq_embed_dim = vq_embed_dim
        g = norm_num_groups
        self.cond_encoder = RefCondEncoder(in_channels, latent_channels, ch, layers_per_block, g,
                                           max_att_resolution, resolution, context_length)
        self.encoder = RefEncoder(in_channels, latent_channels, ch, layers_per_block, g, mid_block_add_attention)
        self.quant_conv = nn.Conv2d(latent_channels, vq_embed_dim, 1)
        self.quantize = RefVectorQuantizer(num_vq_embeddings, vq_embed_dim)
        self.post_quant_conv = nn.Conv2d(vq_embed_dim, latent_channels, 1)
        self.quant_linear = nn.Linear(latent_channels * patch_size * patch_size, vq_embed_dim)
        self.dynamics_quantize = RefVectorQuantizer(num_dyn_embeddings, vq_embed_dim)
        self.post_quant_linear = nn.Linear(vq_embed_dim, latent_channels * patch_size * patch_size)
        self.cond_decoder = RefCondDecoder(latent_channels, out_channels, ch, layers_per_block, g,
                                           max_att_resolution, 16, context_length)
        self.decoder = RefDecoder(latent_channels, out_channels, ch, layers_per_block, g, mid_block_add_attention)

    @classmethod
    def from_config_file(cls, path):
        with open(path) as f:
            cfg = json.load(f)
        cfg = {k: v for k, v in cfg.items() if not k.startswith("_")}
        cfg.pop("down_block_types", None), cfg.pop("up_block_types", None)
        return cls(**cfg)

    # ---- helpers --------------------------------------------------------------------
    def _expand(self, feats: List[torch.Tensor], B: int, fut: int):
        t = self.context_length
        if t > 1:
            return [f.reshape(B, t, *f.shape[-3:]).unsqueeze(1).repeat(1, fut, 1, 1, 1, 1)
                    .reshape(-1, t, *f.shape[-3:]) for f in feats]
        return [f.unsqueeze(1).repeat(1, fut, 1, 1, 1).reshape(-1, *f.shape[-3:]) for f in feats]

    @torch.no_grad()
    def encode_latents(self, pixel_values):
        """Returns pre-quantisation latents (z_ctx [B*t*256,64], z_dyn [B*fut*16,64])."""
        t = self.context_length
        B, T, C, H, W = pixel_values.shape
        fut = T - t
        ctx = pixel_values[:, :t].reshape(-1, C, H, W)
        future = pixel_values[:, t:].reshape(-1, C, H, W)
        h, feats = self.encoder(ctx, return_features=True)
        h = self.quant_conv(h)
        d = self.cond_encoder(future, self._expand(feats, B, fut))
        p = self.patch_size
        d = d.permute(0, 2, 3, 1).unfold(1, p, p).unfold(2, p, p).permute(0, 1, 2, 4, 5, 3)
        d = d.reshape(d.shape[0], d.shape[1] * d.shape[2], -1)
        d = self.quant_linear(d)
        zc = h.permute(0, 2, 3, 1).reshape(-1, self.vq_embed_dim)
        zd = d.reshape(-1, self.vq_embed_dim)
        return zc, zd

    def forward_train(self, sample, dyn_sample, segment_len, idx_ctx=None, idx_dyn=None):
        """compressive_vq_model.py:332-369 (forward) + :290-330 (decode): the tokenizer training graph, differentiable.
        sample [B*t,3,H,W], dyn_sample [B*segment_len,3,H,W] -> (dec, ref_dec, commit_loss, dyn_commit_loss).
        idx_ctx / idx_dyn: see RefVectorQuantizer.forward (test aid).
        Call it in eval(): RefCrossAttention restates the reference block with its dropouts as identity, but the
        nn.MultiheadAttention inside still carries dropout=0.1 and would randomise values and gradients in train()."""
        B = dyn_sample.shape[0] // segment_len
        h, feats = self.encoder(sample, return_features=True)               # :340
        h = self.quant_conv(h)                                              # :349
        d = self.cond_encoder(dyn_sample, self._expand(feats, B, segment_len))      # :351
        p = self.patch_size
        d = d.permute(0, 2, 3, 1).unfold(1, p, p).unfold(2, p, p).permute(0, 1, 2, 4, 5, 3)
        d = d.reshape(d.shape[0], d.shape[1] * d.shape[2], -1)
        d = self.quant_linear(d)                                            # :355
        quant, commit, _ = self.quantize(h, idx=idx_ctx)                    # decode() :297
        dq = d.transpose(-1, -2).unsqueeze(-1)                              # [B, L, D] -> [B, D, L, 1]   :299
        quant_d, dyn_commit, _ = self.dynamics_quantize(dq, idx=idx_dyn)
        quant_d = quant_d.squeeze(-1).transpose(-1, -2)
        q2 = self.post_quant_conv(quant)
        q2d = self.post_quant_linear(quant_d)
        hh, c = q2.shape[-1], self.latent_channels
        q2d = q2d.reshape(q2d.shape[0], hh // p, hh // p, p, p, c)
        q2d = torch.einsum("nhwpqc->nchpwq", q2d).reshape(q2d.shape[0], c, hh, hh)
        ref_dec, dfeats = self.decoder(q2, return_features=True)            # :311
        dec = self.cond_decoder(q2d, self._expand(dfeats, B, segment_len))  # :321
        return dec, ref_dec, commit, dyn_commit

    @torch.no_grad()
    def tokenize(self, pixel_values, context_length=0):
        assert context_length == self.context_length
        t = self.context_length
        B, T = pixel_values.shape[:2]
        fut = T - t
        zc, zd = self.encode_latents(pixel_values)
        ic = torch.argmin(torch.cdist(zc, self.quantize.embedding.weight), dim=1).reshape(B, t, -1)
        idd = torch.argmin(torch.cdist(zd, self.dynamics_quantize.embedding.weight), dim=1).reshape(B, fut, -1)
        return self.serialise(ic, idd)

    def serialise(self, ic, idd):
        """compressive_vq_model.py:205-220."""
        B, t, _ = ic.shape
        fut = idd.shape[1]
        scf = self.num_vq_embeddings + self.num_dyn_embeddings
        sdf = scf + 1
        ic = torch.cat([torch.full((B, t, 1), scf, dtype=ic.dtype), ic], dim=2).reshape(B, -1)[:, 1:]
        idd = torch.cat([torch.full((B, fut, 1), sdf, dtype=idd.dtype), idd + self.num_vq_embeddings], dim=2).reshape(B, -1)
        indices = torch.cat([ic, idd], dim=1)
        labels = torch.cat([torch.full((B, ic.shape[1] + 1), -100, dtype=indices.dtype), idd[:, 1:]], dim=1)
        return indices, labels

    @torch.no_grad()
    def detokenize(self, indices, context_length=0):
        assert context_length == self.context_length
        t, cr, dr = self.context_length, 16, 4
        B = indices.shape[0]
        fut = (indices.shape[1] + 1 - (1 + cr * cr) * t) // (1 + dr * dr)
        idx = torch.cat([torch.ones(B, 1, dtype=indices.dtype), indices], dim=1)
        nct = t * (1 + cr * cr)
        ic = idx[:, :nct].reshape(B, t, -1)[:, :, 1:].reshape(B, -1)
        idd = idx[:, nct:].reshape(B, fut, -1)[:, :, 1:].reshape(B, -1)
        idd = (idd - self.num_vq_embeddings).clamp(0, self.num_dyn_embeddings - 1)
        q = self.quantize.embedding(ic).reshape(B * t, cr, cr, self.vq_embed_dim).permute(0, 3, 1, 2)
        q2 = self.post_quant_conv(q)
        qd = self.dynamics_quantize.embedding(idd).reshape(-1, dr * dr, self.vq_embed_dim)
        q2d = self.post_quant_linear(qd)
        p, c = self.patch_size, self.latent_channels
        q2d = q2d.reshape(q2d.shape[0], cr // p, cr // p, p, p, c)
        q2d = torch.einsum("nhwpqc->nchpwq", q2d).reshape(q2d.shape[0], c, cr, cr)
        ctx_dec, feats = self.decoder(q2, return_features=True)
        dec = self.cond_decoder(q2d, self._expand(feats, B, fut))
        ctx_dec = ctx_dec.reshape(B, t, *ctx_dec.shape[-3:])
        dec = dec.reshape(B, fut, *dec.shape[-3:])
        return torch.cat([ctx_dec, dec], dim=1)


# --------------------------------------------------------------------------------------
# seeded, well-conditioned random weights (no checkpoints exist offline)
# --------------------------------------------------------------------------------------
def seeded_init_(model: nn.Module, seed: int = 1234, codebook: str = "normal") -> nn.Module:
    """Deterministic weights in the reference layout.  torch default init for conv/linear,
    GroupNorm gamma ~ 1+0.1 N(0,1), beta ~ 0.1 N(0,1), small random positional embeddings (the
    reference zero-inits them, which would hide indexing bugs), codebooks N(0,1)*0.5 ("normal")
    or the diffusers U(-1/K,1/K) init ("uniform")."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.named_parameters()):
            if name.endswith("embedding.weight"):
                if codebook == "normal":
                    p.copy_(torch.randn(p.shape, generator=g) * 0.5)
                else:
                    p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / p.shape[0])
            elif "pos_emb" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            elif ("norm" in name) and name.endswith("weight") and p.dim() == 1:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif ("norm" in name) and name.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                bound = 1.0 / math.sqrt(fan_in)
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * bound * math.sqrt(3.0))
            else:
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) * 0.05)
    return model


TINY_CFG = dict(  # a structurally complete miniature (shortcut convs, stride-2, cross-attention in both the
    # encoder and the decoder, head_dim 64) used for fast CPU/GPU parity tests
    in_channels=3, out_channels=3, block_out_channels=(128, 256, 256), layers_per_block=1,
    latent_channels=64, num_vq_embeddings=512, norm_num_groups=32, mid_block_add_attention=False,
    num_dyn_embeddings=512, context_length=2, max_att_resolution=16, resolution=64, patch_size=4)


def config_path(name: str) -> str:
    """Configs are committed copies of plain hyper-parameter JSON under oracle/configs."""
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "configs", name + ".json")
