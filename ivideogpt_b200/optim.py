"""FusedAdamW -- the optimizer step of reference train_gpt.py:803 (`torch.optim.AdamW`, :640-646) as one B200 kernel
launch per parameter (ivgpt_adamw: decoupled weight decay, bias correction, optional gradient scale that carries the
1/world of the data-parallel mean).

It is a regular torch.optim.Optimizer (param groups, state_dict, zero_grad, lr schedulers work unchanged).  The update is
written by a raw-pointer kernel, so every parameter's version counter is bumped explicitly (ops.adamw ->
torch.autograd.graph.increment_version): the packed kernel-layout weight copies of the B200 engines are keyed on
(data_ptr, _version) and MUST be rebuilt after a step -- forgetting this trains on stale matrices.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale: float = 1.0):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError("FusedAdamW: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = float(grad_scale)      # multiplies every gradient (1/world when the exchange was a SUM)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW updates fp32 CUDA parameters only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adamw(p, g.to(torch.float32), st["exp_avg"], st["exp_avg_sq"], float(group["lr"]), b1, b2, group["eps"],
                          group["weight_decay"], st["step"], self.grad_scale)
        return loss
