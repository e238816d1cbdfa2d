"""HeadModelWithAction -- action-conditioned wrapper with the constructor / attribute / method surface of
reference ivideogpt/transformer/action_model.py:8-205, driven by the B200 Llama engine.

Token layout (context c, segment s, 16 tokens per future frame):
    [c0 .. scf .. c_{c-1}] sdf d0[16] sdf d1[16] ...       prelude_tokens_num = 257*c - 1
The embedding of action a_{i+c-1} is added to the input embedding at the i-th sdf slot
(prelude_tokens_num + 17*i), reference :80-81 (generate) and :171-177 (forward).

Status: `generate` / `generate_without_action` / `forward` (loss evaluation) run on the sm_100a kernels.
The reference re-prefills the whole history for every future frame (:78-114, O(frames x history)); here the rollout keeps
ONE KV cache: one prefill of the prompt, then a single decode run (the persistent
megakernel in bf16, the CUDA-graph step in TF32) in which the separator slots are forced to `token_for_sdf` and receive
their action embedding inside the kernel (`slot_cfg` of LlamaEngine.generate).  Same token sequence as the reference's
loop under greedy decoding (tests/test_action_model.py, against vectors produced by the reference file itself);
`persistent_cache = False` selects the reference-shaped per-frame re-prefill loop.  The tiny action/reward linears
(action_dim -> hidden, hidden -> 1) are evaluated with torch.nn.functional.linear on the device: they are not on the
frames/s hot path.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops


def _embed_rows(ids: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    B, L = ids.shape
    out = torch.empty(B, L, table.shape[1], dtype=torch.float32, device=ids.device)   # not a view: callers add to it in place
    ops.embed(ids, ids.stride(0), L, None, table.detach().float().contiguous(), out, B * L)
    return out


class _EmbedFn(torch.autograd.Function):
    """Embedding lookup on the B200 kernels with a gradient for the table (scatter-add of the row gradients)."""

    @staticmethod
    def forward(ctx, ids, table):
        ctx.save_for_backward(ids)
        ctx.shape = tuple(table.shape)
        ctx.tdtype = table.dtype
        return _embed_rows(ids, table)

    @staticmethod
    def backward(ctx, gout):
        (ids,) = ctx.saved_tensors
        dE = torch.zeros(ctx.shape, dtype=torch.float32, device=gout.device)
        g = gout.to(torch.float32).contiguous().view(-1, ctx.shape[1])
        ops.embed_bwd(ids, g, dE)
        return None, dE.to(ctx.tdtype)


class HeadModelWithAction(nn.Module):
    def __init__(self, llm, action_dim, prelude_tokens_num, tokens_num_per_dyna, context, segment_length,
                 model_type='llama', reward_prediction=False, action_recon=None, **kwargs):
        super().__init__()
        if model_type != 'llama':
            raise ValueError(f"model_type {model_type} is not supported by ivideogpt_b200 (llama only).")
        self.llm = llm
        self.action_dim = action_dim
        self.prelude_tokens_num = prelude_tokens_num
        self.tokens_num_per_dyna = tokens_num_per_dyna
        self.context = context
        self.segment_length = segment_length
        self.model_type = model_type
        self.token_for_sdf = llm.config.vocab_size - 1
        self.reward_prediction = reward_prediction
        self.action_recon = action_recon
        hidden = llm.config.hidden_size
        self.action_linear = nn.Linear(action_dim, hidden)
        nn.init.zeros_(self.action_linear.weight)
        nn.init.zeros_(self.action_linear.bias)
        if reward_prediction:
            self.reward_linear = nn.Linear(hidden, 1)
        if action_recon:
            self.action_recon_linear = nn.Linear(hidden, action_dim)
        self.persistent_cache = True          # one KV cache per rollout instead of the reference's per-frame re-prefill

    def _persistent_layout(self, T, per_frame, with_action):
        """(slot0, period) of the forced separator positions when the rollout can keep one cache, else None.
        The reference appends a separator after every `per_frame` generated tokens (:109-110): positions
        T + per_frame + j*(per_frame + 1).  With actions these must coincide with the action slots P + i*(n + 1)."""
        if not self.persistent_cache or not hasattr(self.llm, "b200_engine"):
            return None
        first = T + per_frame
        if not with_action:
            return first, per_frame + 1
        period = self.tokens_num_per_dyna + 1
        if per_frame != self.tokens_num_per_dyna or first < self.prelude_tokens_num or \
                (first - self.prelude_tokens_num) % period != 0 or T <= self.prelude_tokens_num:
            return None
        return self.prelude_tokens_num, period

    # ------------------------------------------------------------------------------------------------------
    def get_input_embeddings(self, input_ids):
        """fp32 token embeddings [B, L, hidden], gathered by the B200 embed kernel.  Differentiable w.r.t. the embedding
        table when autograd is recording (training from inputs_embeds, reference :160-186): the backward is the
        scatter-add kernel ivgpt_embed_bwd."""
        table = self.llm.get_input_embeddings().weight
        if not input_ids.is_cuda:
            raise RuntimeError("HeadModelWithAction requires CUDA tensors (no CPU fallback)")
        ids = input_ids.to(torch.int64).contiguous()
        if torch.is_grad_enabled() and table.requires_grad:
            return _EmbedFn.apply(ids, table)
        return _embed_rows(ids, table)

    def _frames(self):
        return self.segment_length - self.context

    @torch.no_grad()
    @ops.device_scoped
    def generate(self, inputs_token, do_sample=True, temperature=1.0, top_k=100, max_new_tokens=None,
                 pad_token_id=50256, action: Optional[torch.FloatTensor] = None):
        if self.reward_prediction:
            raise NotImplementedError("reward_prediction during generate (reference :83-97, marked buggy there) "
                                      "is not implemented; see mbrl/video_predictor.py for the supported recipe")
        per_frame = ((max_new_tokens + 1) // self._frames()) - 1
        B, T = inputs_token.shape
        act = torch.nn.functional.linear(action.float(), self.action_linear.weight.float(),
                                         self.action_linear.bias.float())
        embeds = self.get_input_embeddings(inputs_token)
        tokens = inputs_token.to(torch.int64)
        layout = self._persistent_layout(T, per_frame, True)
        if layout is not None:
            slot0, period = layout
            n_prompt_slots = (T - 1 - slot0) // period + 1            # slots already inside the prompt (normally 1)
            for i in range(n_prompt_slots):
                embeds[:, slot0 + i * period, :] += act[:, i + self.context - 1, :]
            nslots = n_prompt_slots + self._frames()
            # the reference indexes action[:, i + context - 1] for every generated frame (:80-81) and raises IndexError when
            # the action tensor is too short; only the slot after the LAST frame (forced separator, dropped) may be missing
            need = self.context - 1 + n_prompt_slots - 1 + self._frames()
            if act.shape[1] < need:
                raise IndexError(f"HeadModelWithAction.generate: action has {act.shape[1]} timesteps, the rollout of "
                                 f"{self._frames()} frames with context {self.context} needs at least {need}")
            slot_emb = torch.zeros(B, nslots, act.shape[-1], dtype=torch.float32, device=act.device)
            avail = min(nslots, act.shape[1] - (self.context - 1))
            slot_emb[:, :avail] = act[:, self.context - 1: self.context - 1 + avail]
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if do_sample else 0
            out = self.llm.b200_engine().generate(None, embeds, int(max_new_tokens), bool(do_sample), int(top_k or 0),
                                                  float(temperature), seed,
                                                  slot_cfg=(slot0, period, self.token_for_sdf, slot_emb))
            out[:, :T] = tokens
            return out
        sdf = torch.full((B, 1), self.token_for_sdf, dtype=torch.int64, device=tokens.device)
        for i in range(self._frames()):
            slot = self.prelude_tokens_num + i * (self.tokens_num_per_dyna + 1)
            embeds[:, slot, :] += act[:, i + self.context - 1, :]
            new = self.llm.generate(inputs_embeds=embeds, do_sample=do_sample, temperature=temperature,
                                    top_k=top_k, max_new_tokens=per_frame, pad_token_id=pad_token_id)
            new = torch.cat([new, sdf], dim=1)
            embeds = torch.cat([embeds, self.get_input_embeddings(new)], dim=1)
            tokens = torch.cat([tokens, new], dim=1)
        assert tokens.size(1) == T + max_new_tokens + 1
        return tokens[:, :-1]

    @torch.no_grad()
    @ops.device_scoped
    def generate_without_action(self, inputs_token, do_sample=True, temperature=1.0, top_k=100,
                                max_new_tokens=None):
        per_frame = ((max_new_tokens + 1) // self._frames()) - 1
        B, T = inputs_token.shape
        tokens = inputs_token.to(torch.int64)
        layout = self._persistent_layout(T, per_frame, False)
        if layout is not None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if do_sample else 0
            return self.llm.b200_engine().generate(tokens.contiguous(), None, int(max_new_tokens), bool(do_sample),
                                                   int(top_k or 0), float(temperature), seed,
                                                   slot_cfg=(layout[0], layout[1], self.token_for_sdf, None))
        sdf = torch.full((B, 1), self.token_for_sdf, dtype=torch.int64, device=tokens.device)
        for _ in range(self._frames()):
            new = self.llm.generate(inputs_embeds=self.get_input_embeddings(tokens), do_sample=do_sample,
                                    temperature=temperature, top_k=top_k, max_new_tokens=per_frame)
            tokens = torch.cat([tokens, new, sdf], dim=1)
        assert tokens.size(1) == T + max_new_tokens + 1
        return tokens[:, :-1]

    @ops.device_scoped
    def forward(self, input_ids=None, attention_mask=None, labels=None, position_ids=None, action=None):
        embeds = self.get_input_embeddings(input_ids)
        act = torch.nn.functional.linear(action.float(), self.action_linear.weight.float(),
                                         self.action_linear.bias.float())[:, self.context - 1:-1, :]
        slots = self.prelude_tokens_num + torch.arange(self._frames(), device=embeds.device) * (self.tokens_num_per_dyna + 1)
        embeds[:, slots, :] += act
        x = self.llm(input_ids=None, attention_mask=attention_mask, position_ids=position_ids, inputs_embeds=embeds,
                     labels=labels, output_hidden_states=bool(self.reward_prediction or self.action_recon))
        if self.action_recon:
            hs = x.hidden_states[-1][:, self.prelude_tokens_num:]
            rec = torch.nn.functional.linear(hs, self.action_recon_linear.weight, self.action_recon_linear.bias)
            rec = rec.reshape(-1, self._frames(), self.tokens_num_per_dyna + 1, self.action_dim)
            tgt = action[:, self.context - 1:-1].unsqueeze(-2).repeat(1, 1, self.tokens_num_per_dyna + 1, 1)
            self.action_recon_loss = nn.functional.mse_loss(rec, tgt)
            x.loss = x.loss + self.action_recon * self.action_recon_loss
        if self.reward_prediction:
            hs = x.hidden_states[-1][:, slots + self.tokens_num_per_dyna, :]
            return x, torch.nn.functional.linear(hs, self.reward_linear.weight, self.reward_linear.bias)
        return x
