"""B200LlamaForCausalLM -- the action-free transformer the reference obtains from
`AutoModelForCausalLM.from_pretrained(path, subfolder='transformer')` (inference/predict.py:111) or
`AutoModelForCausalLM.from_config(config)` (train_gpt.py:597).

It subclasses transformers' LlamaForCausalLM ONLY to inherit the parameter tree (identical state-dict keys, so
`model.safetensors` loads with strict=True), config handling and save_pretrained.  `forward` and `generate` are
replaced: on CUDA tensors they run the sm_100a kernel plan in engine.py; on anything else they raise.  No code
path of this class executes transformers' eager/SDPA arithmetic.

Importing `ivideogpt_b200.transformer` registers the class for LlamaConfig with AutoModelForCausalLM -- the
drivers import `ivideogpt.transformer` before they build the model (predict.py:12, train_gpt.py:43), which is
the seam that makes the swap a drop-in.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from transformers import AutoModelForCausalLM, LlamaConfig, LlamaForCausalLM
from transformers.modeling_outputs import CausalLMOutputWithPast

from .. import ops
from .engine import LlamaEngine, LlamaWeights


def _compute_dtype(param_dtype: torch.dtype) -> torch.dtype:
    env = os.environ.get("IVGPT_COMPUTE_DTYPE", "").lower()
    if env in ("bf16", "bfloat16"):
        return torch.bfloat16
    if env in ("fp32", "tf32", "float32"):
        return torch.float32
    if param_dtype in (torch.bfloat16, torch.float16) or torch.is_autocast_enabled():
        return torch.bfloat16
    return torch.float32


class B200LlamaForCausalLM(LlamaForCausalLM):
    _b200_engine: Optional[LlamaEngine] = None
    _b200_sig = None
    _b200_dtype: Optional[torch.dtype] = None

    # ---- engine management ------------------------------------------------------------------------------
    def set_compute_dtype(self, dtype: Optional[torch.dtype]):
        assert dtype in (None, torch.float32, torch.bfloat16)
        self._b200_dtype = dtype
        self._b200_engine = None
        return self

    def b200_engine(self) -> LlamaEngine:
        p = self.lm_head.weight
        if not p.is_cuda:
            raise RuntimeError("B200LlamaForCausalLM runs on a CUDA (sm_100a) device only; move the model with "
                               ".to('cuda').  There is no CPU fallback (use transformers.LlamaForCausalLM for that).")
        dt = self._b200_dtype or _compute_dtype(p.dtype)
        sig = LlamaWeights.signature(self, dt)
        if self._b200_engine is None or self._b200_sig != sig:
            self._b200_engine = LlamaEngine(LlamaWeights(self, dt))
            self._b200_sig = sig
        return self._b200_engine

    @staticmethod
    def _reject(name, value):
        if value is not None and value is not False:
            raise NotImplementedError(f"B200LlamaForCausalLM: argument `{name}` is not supported by the sm_100a path")

    # ---- forward: full-sequence logits (+ shifted CE loss) ---------------------------------------------------------
    @ops.device_scoped
    def forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                inputs_embeds=None, labels=None, use_cache=None, output_attentions=None,
                output_hidden_states=None, return_dict=None, cache_position=None, logits_to_keep=0, **kwargs):
        self._reject("past_key_values", past_key_values)
        self._reject("output_attentions", output_attentions)
        src = input_ids if input_ids is not None else inputs_embeds
        if src is None:
            raise ValueError("either input_ids or inputs_embeds is required")
        if not src.is_cuda:
            raise RuntimeError("B200LlamaForCausalLM.forward: inputs must be CUDA tensors (no CPU fallback)")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            raise NotImplementedError("padding masks are not used on the iVideoGPT path and are not supported")
        B, L = src.shape[0], src.shape[1]
        if position_ids is not None:
            expect = torch.arange(L, device=src.device).expand(B, L)
            if position_ids.shape != expect.shape or not bool((position_ids == expect).all()):
                raise NotImplementedError("only default position_ids (0..L-1) are supported")
        if torch.is_grad_enabled() and labels is not None and any(p.requires_grad for p in self.parameters()):
            # training step (train_gpt.py:792-798): forward + backward on the B200 kernels; the returned loss carries a
            # grad_fn whose backward hands the already-computed parameter gradients to autograd (DDP hooks included).
            if output_hidden_states:
                raise NotImplementedError("training with output_hidden_states (reward / action-reconstruction heads of "
                                          "HeadModelWithAction) is not implemented on the sm_100a path")
            p_drop = float(getattr(self.config, "attention_dropout", 0.0) or 0.0) if self.training else 0.0
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if p_drop > 0.0 else 0      # torch's CPU generator: seedable
            names, params = zip(*self.named_parameters())
            emb = inputs_embeds if input_ids is None else None
            loss = _TrainLossFn.apply(self, input_ids, emb, labels, names, p_drop, seed, *params)
            return CausalLMOutputWithPast(loss=loss, logits=None, past_key_values=None, hidden_states=None,
                                          attentions=None)
        eng = self.b200_engine()
        with torch.no_grad():
            Lmax = (L + 7) // 8 * 8
            want_h = bool(output_hidden_states) or bool(getattr(self.config, "output_hidden_states", False))
            logits_pad, hidden = eng.prefill(B, L, Lmax, input_ids.contiguous() if input_ids is not None else None,
                                             inputs_embeds if input_ids is None else None, True, want_hidden=want_h)
            V = self.config.vocab_size
            loss = None
            if labels is not None:
                loss, _, _ = ops.ce_loss(logits_pad, logits_pad.stride(1), B, L, V, labels.contiguous())
            logits = logits_pad[:, :, :V]
            hs = None
            if want_h:
                hs = (hidden.to(torch.float32).clone(),)   # final-norm output, what callers index with [-1]
        return CausalLMOutputWithPast(loss=loss, logits=logits, past_key_values=None, hidden_states=hs,
                                      attentions=None)

    # ---- generate: prefill + CUDA-graph decode loop -----------------------------------------------------------------
    @torch.no_grad()
    @ops.device_scoped
    def generate(self, inputs=None, generation_config=None, do_sample=None, temperature=None, top_k=None,
                 top_p=None, max_new_tokens=None, max_length=None, pad_token_id=None, eos_token_id=None,
                 use_cache=True, inputs_embeds=None, input_ids=None, attention_mask=None,
                 return_dict_in_generate=False, output_hidden_states=False, num_beams=1, seed=None, **kwargs):
        if inputs is None:
            inputs = input_ids
        if output_hidden_states and not return_dict_in_generate:
            raise NotImplementedError("generate(output_hidden_states=True) needs return_dict_in_generate=True "
                                      "(mbrl/video_predictor.py:293-303 passes both)")
        for _k in ("output_scores", "output_logits", "output_attentions"):
            self._reject(_k, kwargs.get(_k))
        if num_beams not in (None, 1):
            raise NotImplementedError("beam search is not supported")
        if top_p is not None and top_p < 1.0:
            raise NotImplementedError("top_p sampling is not used by iVideoGPT and not supported")
        if attention_mask is not None and not bool((attention_mask != 0).all()):
            raise NotImplementedError("padding masks are not supported")
        gc = self.generation_config
        do_sample = bool(gc.do_sample if do_sample is None else do_sample)
        temperature = float((gc.temperature if gc.temperature is not None else 1.0) if temperature is None else temperature)
        top_k = int((gc.top_k if gc.top_k is not None else 0) if top_k is None else top_k)
        src = inputs if inputs is not None else inputs_embeds
        if src is None:
            raise ValueError("generate needs input ids or inputs_embeds")
        if not src.is_cuda:
            raise RuntimeError("B200LlamaForCausalLM.generate: inputs must be CUDA tensors (no CPU fallback)")
        L = src.shape[1]
        if max_new_tokens is None:
            if max_length is None:
                raise ValueError("max_new_tokens (or max_length) is required")
            max_new_tokens = int(max_length) - L
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if do_sample else 0
        eng = self.b200_engine()
        tokens = eng.generate(inputs.contiguous() if inputs is not None else None,
                              inputs_embeds if inputs is None else None, int(max_new_tokens), do_sample, top_k,
                              temperature, seed)
        # eos_token_id 50256 is outside the 16386-token vocabulary (configs/llama/config.json) -> never stops early.
        # HF returns only the new tokens when fed inputs_embeds (action_model.py:101-114)
        sequences = tokens[:, L:] if inputs is None else tokens
        if not return_dict_in_generate:
            return sequences
        from transformers.generation.utils import GenerateDecoderOnlyOutput
        hidden_states = None
        if output_hidden_states:
            hidden_states = self._generation_hidden_states(eng, inputs, inputs_embeds if inputs is None else None, tokens,
                                                           L, int(max_new_tokens))
        return GenerateDecoderOnlyOutput(sequences=sequences, hidden_states=hidden_states)

    def _generation_hidden_states(self, eng, inputs, inputs_embeds, tokens, L, new):
        """`result.hidden_states` of HF generate(return_dict_in_generate=True, output_hidden_states=True), as far as the
        reference reads it (mbrl/video_predictor.py:305-308: hidden_states[-1][-1] -> reward head): a tuple over the `new`
        generation steps, each a tuple whose LAST entry is the final-norm output of the positions fed at that step
        ([B, L, h] for the prompt step, [B, 1, h] afterwards).  HF's tuples also hold the embedding output and every
        intermediate layer; only the last entry is provided here (1-tuples).  The states come from ONE teacher-forced pass
        over prompt + generated tokens on the prefill kernels after the rollout (same quantities as the decode steps up to
        rounding), so the decode kernels carry no per-step hidden-state traffic."""
        B = tokens.shape[0]
        fed = L + new - 1                                    # positions that were fed to the model
        if fed <= 0:
            return ()
        Lmax = (fed + 7) // 8 * 8
        if inputs is not None:
            _, hidden = eng.prefill(B, fed, Lmax, tokens[:, :fed].contiguous(), None, True, hidden_only=True)
        else:
            emb = inputs_embeds.to(torch.float32)
            if new > 1:
                ids_new = tokens[:, L:fed].contiguous()
                table = eng.w.embed
                rows = torch.empty(B * (fed - L), table.shape[1], dtype=torch.float32, device=tokens.device)
                ops.embed(ids_new, ids_new.stride(0), fed - L, None, table, rows, B * (fed - L))
                emb = torch.cat([emb, rows.view(B, fed - L, -1)], dim=1)
            _, hidden = eng.prefill(B, fed, Lmax, None, emb.contiguous(), True, hidden_only=True)
        hidden = hidden.to(torch.float32).clone()
        steps = [(hidden[:, :L],)] + [(hidden[:, L + i - 1: L + i],) for i in range(1, new)]
        return tuple(steps)

    def gradient_checkpointing_enable(self, *a, **k):  # accepted and ignored (train_gpt.py:598-600)
        return None


class _TrainLossFn(torch.autograd.Function):
    """loss = Llama(input_ids, labels); gradients of every parameter are produced by LlamaTrainEngine in the forward
    call (activations never outlive it) and released to autograd in backward."""

    @staticmethod
    def forward(ctx, model, input_ids, inputs_embeds, labels, names, p_drop, seed, *params):
        from .train_engine import LlamaTrainEngine
        eng = model.b200_engine()
        train = LlamaTrainEngine(eng.w)
        # model.b200_grad_reducer (grad_reduce.BucketedGradReducer, optional): the data-parallel exchange, launched bucket
        # by bucket from inside the backward so that it overlaps the remaining layers (row a13, train_gpt.py:672,798)
        reducer = getattr(model, "b200_grad_reducer", None)
        loss, grads = train.forward_backward(input_ids, labels, on_grads=reducer.on_grads if reducer is not None else None,
                                             embeds=inputs_embeds, attn_dropout=p_drop, seed=seed)
        if reducer is not None:
            reducer.finish(grads)
        # trained from inputs_embeds: the embedding table is reached through the caller's own lookup (autograd), not here
        ctx.grads = [grads.get(n) for n in names]
        ctx.d_embeds = grads.get("inputs_embeds")
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        out = []
        for g in ctx.grads:
            out.append(None if g is None else (g.mul_(gout) if g.is_contiguous() else g * gout))
        d_emb = None if ctx.d_embeds is None else ctx.d_embeds * gout
        ctx.grads = None
        ctx.d_embeds = None
        return (None, None, d_emb, None, None, None, None) + tuple(out)


def register():
    AutoModelForCausalLM.register(LlamaConfig, B200LlamaForCausalLM, exist_ok=True)
