"""B200-native CompressiveVQModel: same Python surface as reference
ivideogpt/vq_model/compressive_vq_model.py (class :33, __init__ :36-152, set_context_length :154,
tokenize :165-220, detokenize :223-277), none of its code.

Differences that matter to a caller:  (1) it runs on CUDA sm_100a only and raises otherwise -- there is no
PyTorch fallback;  (2) `diffusers` is not needed: from_pretrained / from_config / save_pretrained / .config
are implemented in hub_io.py against the same on-disk layout (config.json + diffusion_pytorch_model.safetensors);
(3) `forward()` (the tokenizer training graph, reference :332-369) runs FORWARD ONLY (evaluation, validation losses);
its backward is not built and raises.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import ops
from .hub_io import HubMixin
from .modules import Codebook, DecoderParams, EncoderParams
from .plan import TokenizerPlan

_CTX_RES = 16   # latent side of a context frame  (reference :225 "magic number")
_DYN_RES = 4    # token-grid side of a future frame (reference :226)


def _default_compute_dtype(param_dtype: torch.dtype) -> torch.dtype:
    env = os.environ.get("IVGPT_COMPUTE_DTYPE", "").lower()
    if env in ("bf16", "bfloat16"):
        return torch.bfloat16
    if env in ("fp32", "tf32", "float32"):
        return torch.float32
    if param_dtype == torch.bfloat16 or torch.is_autocast_enabled():
        return torch.bfloat16
    return torch.float32


class CompressiveVQDecoderOutput:
    """Field-for-field the reference's output record (compressive_vq_model.py:16-30)."""

    def __init__(self, sample, ref_sample=None, commit_loss=None, dyn_commit_loss=None):
        self.sample, self.ref_sample, self.commit_loss, self.dyn_commit_loss = sample, ref_sample, commit_loss, dyn_commit_loss


class CompressiveVQModel(HubMixin, nn.Module):
    config_name = "config.json"

    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 3,
        down_block_types: Tuple[str, ...] = ("DownEncoderBlock2D",),
        up_block_types: Tuple[str, ...] = ("UpDecoderBlock2D",),
        block_out_channels: Tuple[int, ...] = (64,),
        layers_per_block: int = 1,
        act_fn: str = "silu",
        latent_channels: int = 3,
        sample_size: int = 32,
        num_vq_embeddings: int = 256,
        norm_num_groups: int = 32,
        vq_embed_dim: Optional[int] = None,
        scaling_factor: float = 0.18215,
        norm_type: str = "group",
        mid_block_add_attention=True,
        lookup_from_codebook=False,
        force_upcast=False,
        num_dyn_embeddings: int = 256,
        context_length: int = 1,
        max_att_resolution=32,
        resolution=256,
        patch_size=4,
    ):
        super().__init__()
        self.register_config(locals())
        if act_fn != "silu" or norm_type != "group":
            raise NotImplementedError("ivideogpt_b200 tokenizer kernels implement act_fn='silu', norm_type='group'")
        if any(t != "DownEncoderBlock2D" for t in down_block_types) or \
                any(t != "UpDecoderBlock2D" for t in up_block_types):
            raise NotImplementedError("only DownEncoderBlock2D / UpDecoderBlock2D stages are supported")
        chans = tuple(block_out_channels)
        self.latent_channels = latent_channels
        self.dyna_latent_channels = latent_channels
        self.context_length = context_length
        self.num_vq_embeddings = num_vq_embeddings
        self.num_dyn_embeddings = num_dyn_embeddings
        self.patch_size = patch_size
        vq_embed_dim = vq_embed_dim if vq_embed_dim is not None else latent_channels
        self.vq_embed_dim = vq_embed_dim
        g = norm_num_groups

        self.cond_encoder = EncoderParams(in_channels, latent_channels, chans, layers_per_block, g, True, cross=True,
                                          max_att=max_att_resolution, init_res=resolution, ctx=context_length)
        self.encoder = EncoderParams(in_channels, latent_channels, chans, layers_per_block, g, mid_block_add_attention)
        self.quant_conv = nn.Conv2d(latent_channels, vq_embed_dim, 1)
        self.quantize = Codebook(num_vq_embeddings, vq_embed_dim)
        self.post_quant_conv = nn.Conv2d(vq_embed_dim, latent_channels, 1)
        self.quant_linear = nn.Linear(latent_channels * patch_size * patch_size, vq_embed_dim)
        self.dynamics_quantize = Codebook(num_dyn_embeddings, vq_embed_dim)
        self.post_quant_linear = nn.Linear(vq_embed_dim, latent_channels * patch_size * patch_size)
        self.cond_decoder = DecoderParams(latent_channels, out_channels, chans, layers_per_block, g, True, cross=True,
                                          max_att=max_att_resolution, init_res=_CTX_RES, ctx=context_length)
        self.decoder = DecoderParams(latent_channels, out_channels, chans, layers_per_block, g, mid_block_add_attention)

        self._plan: Optional[TokenizerPlan] = None
        self._compute_dtype: Optional[torch.dtype] = None
        self._groups = g

    # ---- reference-compatible small API -----------------------------------------------------------
    def set_context_length(self, context_length):
        self.context_length = context_length
        self.config["context_length"] = context_length
        self.cond_encoder.set_context_length(context_length)
        self.cond_decoder.set_context_length(context_length)
        if self._plan is not None:
            self._plan.pw.clear()

    def init_modules(self):
        print(self.cond_decoder.load_state_dict(self.decoder.state_dict(), strict=False))
        print(self.cond_encoder.load_state_dict(self.encoder.state_dict(), strict=False))

    def set_compute_dtype(self, dtype: Optional[torch.dtype]):
        """torch.float32 (TF32 tensor cores, fp32 storage), torch.bfloat16, or None (= infer per call)."""
        assert dtype in (None, torch.float32, torch.bfloat16)
        self._compute_dtype = dtype
        self._plan = None
        return self

    def _get_plan(self) -> TokenizerPlan:
        dt = self._compute_dtype or _default_compute_dtype(self.quant_conv.weight.dtype)
        if self._plan is None or self._plan.dtype != dt:
            self._plan = TokenizerPlan(self._groups, dt)
        return self._plan

    def _require_cuda(self, t: torch.Tensor, what: str):
        if not t.is_cuda or not self.quant_conv.weight.is_cuda:
            raise RuntimeError(f"ivideogpt_b200.CompressiveVQModel.{what}: model and inputs must be on a CUDA "
                               f"(sm_100a) device; got input on {t.device}, weights on "
                               f"{self.quant_conv.weight.device}.  There is no CPU fallback.")

    # ---- tokenize -----------------------------------------------------------------------------------
    @torch.no_grad()
    @ops.device_scoped
    def encode_latents(self, pixel_values: torch.Tensor, with_dynamics: bool = True):
        """Pre-quantisation latents in fp32: z_ctx [B*t*256, D], z_dyn [B*f*16, D] (None if not requested)."""
        self._require_cuda(pixel_values, "tokenize")
        plan = self._get_plan()
        t = self.context_length
        B, T, Cc, H, W = pixel_values.shape
        f = T - t
        px = pixel_values.to(torch.float32).contiguous()
        h, feats = plan.encode(px, self.encoder, 0, t, want_features=True)
        wq, bq = plan.pw.linear(self.quant_conv.weight, self.quant_conv.bias, plan.dtype)
        z_ctx = ops.gemm(h.view(-1, h.shape[-1]), wq, bq, out_dtype=torch.float32)
        z_dyn = None
        if with_dynamics and f > 0:
            d = plan.encode(px, self.cond_encoder, t, f, ctx_feats=feats)
            patches = ops.patchify(d, self.patch_size)
            wl, bl = plan.pw.linear(self.quant_linear.weight, self.quant_linear.bias, plan.dtype)
            z_dyn = ops.gemm(patches, wl, bl, out_dtype=torch.float32)
        return z_ctx, z_dyn

    @torch.no_grad()
    @ops.device_scoped
    def tokenize(self, pixel_values: torch.FloatTensor, context_length: int = 0):
        assert context_length == self.context_length  # same contract as the reference (:166)
        t = self.context_length
        B, T = pixel_values.shape[:2]
        f = T - t
        z_ctx, z_dyn = self.encode_latents(pixel_values)
        idx_c = ops.vq_argmin(z_ctx, self.quantize.embedding.weight.detach().float())
        if f > 0:
            idx_d = ops.vq_argmin(z_dyn, self.dynamics_quantize.embedding.weight.detach().float())
        else:
            idx_d = torch.empty(0, dtype=torch.int64, device=idx_c.device)
        tokens, labels = ops.tokens_serialise(idx_c, idx_d, B, t, f, _CTX_RES * _CTX_RES, _DYN_RES * _DYN_RES,
                                              self.num_vq_embeddings, self.num_dyn_embeddings)
        return tokens, labels

    @torch.no_grad()
    @ops.device_scoped
    def tokenize_context(self, pixel_values: torch.FloatTensor):
        """Prediction-minimal entry: tokens of the context frames only (what predict.py:54 keeps)."""
        t = self.context_length
        B = pixel_values.shape[0]
        z_ctx, _ = self.encode_latents(pixel_values[:, :t], with_dynamics=False)
        idx_c = ops.vq_argmin(z_ctx, self.quantize.embedding.weight.detach().float())
        # serialise with one placeholder future frame so that the trailing sdf separator (:211-213) is emitted,
        # then keep the t*257 prompt tokens -- exactly `tokens[:, :context_length * 257]` of predict.py:54.
        idx_d = torch.zeros(B * _DYN_RES * _DYN_RES, dtype=torch.int64, device=idx_c.device)
        tokens, _ = ops.tokens_serialise(idx_c, idx_d, B, t, 1, _CTX_RES * _CTX_RES, _DYN_RES * _DYN_RES,
                                         self.num_vq_embeddings, self.num_dyn_embeddings, want_labels=False)
        return tokens[:, : t * (_CTX_RES * _CTX_RES + 1)]

    # ---- detokenize ---------------------------------------------------------------------------------
    @torch.no_grad()
    @ops.device_scoped
    def detokenize(self, indices, context_length: int = 0, cache=None, return_cache=False):
        assert context_length == self.context_length
        self._require_cuda(indices, "detokenize")
        plan = self._get_plan()
        t, cr, dr = self.context_length, _CTX_RES, _DYN_RES
        B, L = indices.shape
        assert (L + 1 - (1 + cr * cr) * t) % (1 + dr * dr) == 0
        f = (L + 1 - (1 + cr * cr) * t) // (1 + dr * dr)
        D = self.vq_embed_dim
        bad_ctx = torch.zeros(1, dtype=torch.int32, device=indices.device)
        qc, qd = ops.tokens_gather(indices.contiguous(), self.quantize.embedding.weight.detach().float(),
                                   self.dynamics_quantize.embedding.weight.detach().float(), t, f, cr * cr, dr * dr,
                                   plan.dtype, bad_ctx=bad_ctx)
        H = W = self.config["resolution"] if "resolution" in self.config else None
        out = None
        if cache is not None:
            ctx_feats, ctx_frames = cache["cond_features"], cache["context_dec"]
            H, W = ctx_frames.shape[-2:]
            out = torch.empty(B, t + f, self.config["out_channels"], H, W, dtype=torch.float32, device=indices.device)
            out[:, :t] = ctx_frames
        else:
            wpc, bpc = plan.pw.linear(self.post_quant_conv.weight, self.post_quant_conv.bias, plan.dtype)
            lat_c = ops.gemm(qc, wpc, bpc).view(B * t, cr, cr, self.latent_channels)
            # output spatial size follows from the number of up-sampling stages
            scale = 2 ** (len(self.decoder.up_blocks) - 1)
            H = W = cr * scale
            out = torch.empty(B, t + f, self.config["out_channels"], H, W, dtype=torch.float32, device=indices.device)
            ctx_feats = plan.decode(lat_c, self.decoder, out, 0, t, want_features=True)
        if f > 0:
            wpl, bpl = plan.pw.linear(self.post_quant_linear.weight, self.post_quant_linear.bias, plan.dtype)
            pd = ops.gemm(qd, wpl, bpl)                                       # [B*f*16, p*p*latent]
            lat_d = ops.patchify(pd, self.patch_size, inverse=True, frames=B * f, res=cr, ch=self.latent_channels)
            plan.decode(lat_d, self.cond_decoder, out, t, f, ctx_feats=ctx_feats)
        # every kernel of the call is enqueued: now look at the flag (one host sync at the END of the call; the reference's
        # embedding lookup raises on such ids, a mis-sliced prompt must not decode silently to wrong frames)
        if int(bad_ctx.item()) != 0:
            raise IndexError("detokenize: a context position holds a token id outside the context codebook "
                             f"[0, {self.num_vq_embeddings}) -- separator or dynamics ids in a context slot "
                             "(reference compressive_vq_model.py:238 raises in the embedding lookup)")
        if return_cache:
            return out, {"context_dec": out[:, :t].clone(), "cond_features": ctx_feats}
        return out

    # ---- forward (tokenizer training / evaluation graph, reference :332-369 + decode() :290-330) ---------------------------
    @ops.device_scoped
    def forward(self, sample: torch.FloatTensor, return_dict: bool = True, return_loss: bool = False,
                segment_len: int = None, dyn_sample: torch.FloatTensor = None):
        """sample [B*t, 3, H, W] context frames, dyn_sample [B*segment_len, 3, H, W] future frames ->
        (dec [B*segment_len, 3, H, W], ref_dec [B*t, 3, H, W], commit_loss, dyn_commit_loss), the values the reference's
        forward returns (encoder -> quant_conv -> VQ (straight-through value, commit loss) -> post_quant_conv -> decoder;
        conditional encoder -> patchify -> quant_linear -> VQ -> post_quant_linear -> de-patchify -> conditional decoder).

        With autograd recording and trainable parameters this is the TRAINING graph (reference train_tokenizer.py:734
        `accelerator.backward(loss)`): the forward keeps a tape and `loss.backward()` runs the conv dgrad / wgrad, GroupNorm,
        attention, codebook and straight-through gradients on the sm_100a kernels (vq_model/train_plan.py; fp32 storage /
        TF32 tensor cores).  Under torch.no_grad(), or with frozen parameters, it is the evaluation forward of
        train_tokenizer.py's validation loop (:940-960) in the configured compute dtype."""
        if dyn_sample is None or segment_len is None:
            raise NotImplementedError("CompressiveVQModel.forward: the ctx_vqgan form (sample=, dyn_sample=, segment_len=) "
                                      "is the one train_tokenizer.py uses (:623-627) and the one implemented")
        self._require_cuda(sample, "forward")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .train_plan import forward_train
            t, f = self.context_length, int(segment_len)
            assert sample.shape[0] % t == 0 and dyn_sample.shape[0] % f == 0 and sample.shape[0] // t == dyn_sample.shape[0] // f
            (dec, ref_dec, commit, dyn_commit), self._last_train_graph = forward_train(
                self, self._get_plan(), sample, dyn_sample, f, getattr(self, "_vq_index_override", None))
            if not return_dict:
                return (dec, ref_dec, commit, dyn_commit) if return_loss else (dec,)
            if return_loss:
                return CompressiveVQDecoderOutput(sample=dec, ref_sample=ref_dec, commit_loss=commit, dyn_commit_loss=dyn_commit)
            return CompressiveVQDecoderOutput(sample=dec)
        self._require_cuda(sample, "forward")
        with torch.no_grad():
            plan = self._get_plan()
            t, f, cr, dr = self.context_length, int(segment_len), _CTX_RES, _DYN_RES
            assert sample.shape[0] % t == 0 and dyn_sample.shape[0] % f == 0 and sample.shape[0] // t == dyn_sample.shape[0] // f
            B = sample.shape[0] // t
            Cc, H, W = sample.shape[1:]
            ctx = sample.to(torch.float32).contiguous().view(B, t, Cc, H, W)
            fut = dyn_sample.to(torch.float32).contiguous().view(B, f, Cc, H, W)
            # context branch
            h, feats = plan.encode(ctx, self.encoder, 0, t, want_features=True)
            wq, bq = plan.pw.linear(self.quant_conv.weight, self.quant_conv.bias, plan.dtype)
            z_ctx = ops.gemm(h.view(-1, h.shape[-1]), wq, bq, out_dtype=torch.float32)
            cb_c = self.quantize.embedding.weight.detach().float()
            zq_c, commit = ops.vq_commit(z_ctx, cb_c, ops.vq_argmin(z_ctx, cb_c), plan.dtype, beta=1.0)
            wpc, bpc = plan.pw.linear(self.post_quant_conv.weight, self.post_quant_conv.bias, plan.dtype)
            lat_c = ops.gemm(zq_c, wpc, bpc).view(B * t, cr, cr, self.latent_channels)
            ref_dec = torch.empty(B, t, self.config["out_channels"], H, W, dtype=torch.float32, device=sample.device)
            dec_feats = plan.decode(lat_c, self.decoder, ref_dec, 0, t, want_features=True)
            # dynamics branch
            d = plan.encode(fut, self.cond_encoder, 0, f, ctx_feats=feats)
            patches = ops.patchify(d, self.patch_size)
            wl, bl = plan.pw.linear(self.quant_linear.weight, self.quant_linear.bias, plan.dtype)
            z_dyn = ops.gemm(patches, wl, bl, out_dtype=torch.float32)
            cb_d = self.dynamics_quantize.embedding.weight.detach().float()
            zq_d, dyn_commit = ops.vq_commit(z_dyn, cb_d, ops.vq_argmin(z_dyn, cb_d), plan.dtype, beta=1.0)
            wpl, bpl = plan.pw.linear(self.post_quant_linear.weight, self.post_quant_linear.bias, plan.dtype)
            pd = ops.gemm(zq_d, wpl, bpl)
            lat_d = ops.patchify(pd, self.patch_size, inverse=True, frames=B * f, res=cr, ch=self.latent_channels)
            dec = torch.empty(B, f, self.config["out_channels"], H, W, dtype=torch.float32, device=sample.device)
            plan.decode(lat_d, self.cond_decoder, dec, 0, f, ctx_feats=dec_feats)
            dec = dec.view(B * f, -1, H, W)
            ref_dec = ref_dec.view(B * t, -1, H, W)
        if not return_dict:
            return (dec, ref_dec, commit, dyn_commit) if return_loss else (dec,)
        if return_loss:
            return CompressiveVQDecoderOutput(sample=dec, ref_sample=ref_dec, commit_loss=commit, dyn_commit_loss=dyn_commit)
        return CompressiveVQDecoderOutput(sample=dec)
