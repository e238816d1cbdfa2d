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
            if input_ids is None:
                raise NotImplementedError("training from inputs_embeds (action-conditioned fine-tuning) is not built yet")
            if float(getattr(self.config, "attention_dropout", 0.0) or 0.0) > 0.0 and self.training:
                import warnings
                warnings.warn("B200LlamaForCausalLM: attention_dropout > 0 is ignored (dropout 0 is what loss parity "
                              "is defined on; see DESIGN.md)", stacklevel=2)
            names, params = zip(*self.named_parameters())
            loss = _TrainLossFn.apply(self, input_ids, labels, names, *params)
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
    def generate(self, inputs=None, generation_config=None, do_sample=None, temperature=None, top_k=None,
                 top_p=None, max_new_tokens=None, max_length=None, pad_token_id=None, eos_token_id=None,
                 use_cache=True, inputs_embeds=None, input_ids=None, attention_mask=None,
                 return_dict_in_generate=False, output_hidden_states=False, num_beams=1, seed=None, **kwargs):
        if inputs is None:
            inputs = input_ids
        self._reject("return_dict_in_generate", return_dict_in_generate)
        self._reject("output_hidden_states", output_hidden_states)
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
        if inputs is None:
            return tokens[:, L:]      # HF returns only the new tokens when fed inputs_embeds (action_model.py:101-114)
        return tokens

    def gradient_checkpointing_enable(self, *a, **k):  # accepted and ignored (train_gpt.py:598-600)
        return None


class _TrainLossFn(torch.autograd.Function):
    """loss = Llama(input_ids, labels); gradients of every parameter are produced by LlamaTrainEngine in the forward
    call (activations never outlive it) and released to autograd in backward."""

    @staticmethod
    def forward(ctx, model, input_ids, labels, names, *params):
        from .train_engine import LlamaTrainEngine
        eng = model.b200_engine()
        train = LlamaTrainEngine(eng.w)
        # model.b200_grad_reducer (grad_reduce.BucketedGradReducer, optional): the data-parallel exchange, launched bucket
        # by bucket from inside the backward so that it overlaps the remaining layers (row a13, train_gpt.py:672,798)
        reducer = getattr(model, "b200_grad_reducer", None)
        loss, grads = train.forward_backward(input_ids, labels, on_grads=reducer.on_grads if reducer is not None else None)
        if reducer is not None:
            reducer.finish(grads)
        ctx.grads = [grads[n] for n in names]
        return loss.clone()

    @staticmethod
    def backward(ctx, gout):
        out = []
        for g in ctx.grads:
            out.append(g.mul_(gout) if g.is_contiguous() else g * gout)
        ctx.grads = None
        return (None, None, None, None) + tuple(out)


def register():
    AutoModelForCausalLM.register(LlamaConfig, B200LlamaForCausalLM, exist_ok=True)
