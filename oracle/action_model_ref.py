"""ORACLE (test infrastructure, NOT product code) -- action-conditioned wrapper of the transformer half.

CPU restatement of reference ivideogpt/transformer/action_model.py on top of the UNMODIFIED HF LlamaForCausalLM
(oracle/llama_ref.py).  Pinned: tests/golden/action_model_tiny.npz was produced by importing the reference file itself
in the build container (tests/golden/make_golden_action.py); tests/test_action_model.py checks this restatement
against those vectors on CPU, and the B200 product against both on the GPU.

Sequence layout (reference :11-16, :66-70):  prelude tokens [0, P) | sdf | d[0:n] | sdf | d[0:n] | ...
The i-th sdf slot sits at P + i*(n+1) and receives action_linear(action[:, context-1+i]) on top of its token embedding
(:80-81 generate, :171-177 forward).
"""
from __future__ import annotations

import torch
from torch import nn


class RefActionModel(nn.Module):
    def __init__(self, llm, action_dim, prelude_tokens_num, tokens_num_per_dyna, context, segment_length,
                 reward_prediction=False, action_recon=None):
        super().__init__()
        h = llm.config.hidden_size
        self.llm = llm
        self.P, self.n, self.context, self.segment = prelude_tokens_num, tokens_num_per_dyna, context, segment_length
        self.frames = segment_length - context
        self.sdf = llm.config.vocab_size - 1                      # :26
        self.action_linear = nn.Linear(action_dim, h)             # :35 (zero-initialised at :37-38)
        nn.init.zeros_(self.action_linear.weight), nn.init.zeros_(self.action_linear.bias)
        self.reward_linear = nn.Linear(h, 1) if reward_prediction else None            # :40-41
        self.action_recon = action_recon
        self.action_recon_linear = nn.Linear(h, action_dim) if action_recon else None  # :43-44

    def slots(self):
        return self.P + torch.arange(self.frames) * (self.n + 1)  # :169-170

    def embed(self, ids):
        return self.llm.get_input_embeddings()(ids)               # :46-53

    @torch.no_grad()
    def generate(self, tokens, action, max_new_tokens, do_sample=False, top_k=100, temperature=1.0):
        """:56-121 -- one HF generate per future frame over the whole (growing) embedded history."""
        per_frame = (max_new_tokens + 1) // self.frames - 1       # :72
        act = self.action_linear(action)
        emb = self.embed(tokens)
        T = tokens.shape[1]
        for i in range(self.frames):
            emb[:, self.P + i * (self.n + 1)] += act[:, i + self.context - 1]       # :80-81
            new = self.llm.generate(inputs_embeds=emb, do_sample=do_sample, temperature=temperature, top_k=top_k,
                                    max_new_tokens=per_frame, pad_token_id=50256, use_cache=True)   # :99-108
            new = torch.cat([new, torch.full_like(new[:, :1], self.sdf)], dim=1)    # :109-110
            emb = torch.cat([emb, self.embed(new)], dim=1)
            tokens = torch.cat([tokens, new], dim=1)
        assert tokens.shape[1] == T + max_new_tokens + 1          # :115
        return tokens[:, :-1]                                     # :121 the trailing sdf is dropped

    @torch.no_grad()
    def generate_without_action(self, tokens, max_new_tokens, do_sample=False, top_k=100, temperature=1.0):
        """:123-151."""
        per_frame = (max_new_tokens + 1) // self.frames - 1
        for _ in range(self.frames):
            new = self.llm.generate(inputs_embeds=self.embed(tokens), do_sample=do_sample, temperature=temperature,
                                    top_k=top_k, max_new_tokens=per_frame, pad_token_id=50256)
            tokens = torch.cat([tokens, new, torch.full_like(new[:, :1], self.sdf)], dim=1).to(torch.int64)
        return tokens[:, :-1]

    def forward(self, input_ids, labels, action):
        """:154-205 -- returns (loss, logits, reward_pred or None)."""
        emb = self.embed(input_ids).clone()
        emb[:, self.slots()] += self.action_linear(action)[:, self.context - 1:-1]  # :166-171
        want_h = self.reward_linear is not None or self.action_recon_linear is not None
        out = self.llm(inputs_embeds=emb, labels=labels, output_hidden_states=want_h)
        loss = out.loss
        if self.action_recon_linear is not None:                  # :182-190
            hs = out.hidden_states[-1][:, self.P:]
            rec = self.action_recon_linear(hs).reshape(-1, self.frames, self.n + 1, action.shape[-1])
            tgt = action[:, self.context - 1:-1].unsqueeze(-2).repeat(1, 1, self.n + 1, 1)
            loss = loss + self.action_recon * nn.functional.mse_loss(rec, tgt)
        reward = None
        if self.reward_linear is not None:                        # :192-198
            reward = self.reward_linear(out.hidden_states[-1][:, self.slots() + self.n])
        return loss, out.logits, reward


def seeded_heads_(m: RefActionModel, seed: int = 77, scale: float = 0.5):
    """Non-zero seeded action / reward / recon heads (the zero init would make the action path invisible to a test)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(m.named_parameters()):
            if name.startswith("llm."):
                continue
            p.copy_(torch.randn(p.shape, generator=g) * scale)
    return m
