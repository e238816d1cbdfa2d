"""Data-parallel training exchange: bucketed gradient all-reduce overlapped with the backward pass.

Replaces what accelerate's DistributedDataParallel does for reference train_gpt.py:672,798 (mean of the gradients over
ranks, bucketed, overlapped with backward).  The B200 training engine produces a layer's parameter gradients in one go
(transformer/train_engine.py); as soon as a layer is done its gradients are packed into ONE flat fp32 bucket and an
asynchronous NCCL all-reduce is launched on NCCL's own stream while the engine keeps computing the next (earlier) layer.
`finish()` makes the compute stream wait for the outstanding buckets and re-points every gradient at its slice of the
reduced bucket.  Buckets follow backward order: [lm_head, final norm], layer L-1, ..., layer 0, [embedding] -- the
largest single bucket of the 138M model is the lm_head (50 MB), a layer is 28 MB: sized for launch latency and overlap,
not link count (NVSwitch gives every GPU full bandwidth to every peer).

The sum (not the mean) is exchanged; the 1/world factor is folded into the optimizer's gradient scale (ivgpt_adamw
`gscale`), exactly one multiply per element either way.

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests (tests/test_grad_reduce.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


class BucketedGradReducer:
    def __init__(self, group: Optional[dist.ProcessGroup] = None, min_bucket_bytes: int = 0):
        self.group = group
        self.min_bucket_bytes = int(min_bucket_bytes)     # merge consecutive small buckets (0: one bucket per call)
        self._pending: List[tuple] = []                    # (work, flat, names, shapes)
        self._carry_names: List[str] = []
        self.buckets_launched = 0
        self.bytes_reduced = 0

    @property
    def world(self) -> int:
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def on_grads(self, grads: Dict[str, torch.Tensor], names: Sequence[str]) -> None:
        """Called by the training engine when grads[n] for n in names are final."""
        self._carry_names.extend(names)
        nbytes = sum(grads[n].numel() * grads[n].element_size() for n in self._carry_names)
        if nbytes >= self.min_bucket_bytes:
            self._launch(grads)

    def _launch(self, grads: Dict[str, torch.Tensor]) -> None:
        names, self._carry_names = self._carry_names, []
        if not names:
            return
        shapes = [tuple(grads[n].shape) for n in names]
        flat = torch.cat([grads[n].reshape(-1) for n in names])          # one packed bucket (also makes strided views dense)
        work = None
        if self.world > 1:
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._pending.append((work, flat, names, shapes))
        self.buckets_launched += 1
        self.bytes_reduced += flat.numel() * flat.element_size()
        # drop the unreduced copies now: the dict entries are re-pointed at bucket slices in finish()
        for n in names:
            grads[n] = None

    def finish(self, grads: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Waits for every outstanding bucket (stream-ordered on CUDA) and fills grads with views of the reduced buckets."""
        self._launch(grads)
        for work, flat, names, shapes in self._pending:
            if work is not None:
                work.wait()
            off = 0
            for n, shp in zip(names, shapes):
                cnt = 1
                for d in shp:
                    cnt *= d
                grads[n] = flat[off:off + cnt].view(shp)
                off += cnt
        self._pending = []
        return grads


def allreduce_grads_flat(params, group: Optional[dist.ProcessGroup] = None) -> int:
    """One flat all-reduce (SUM) of every existing `.grad` of `params` after the backward -- the simple exchange used by the
    tokenizer training step (reference train_tokenizer.py:734 under DDP), where one autograd node produces all gradients at
    once and there is nothing to overlap with yet.  Parameters without a gradient are skipped (all ranks must agree on which,
    as under DDP).  Returns the bytes exchanged; a single process is a no-op."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    gs = [p.grad for p in params if p.grad is not None]
    if not gs:
        return 0
    flat = torch.cat([g.reshape(-1) for g in gs])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in gs:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel() * flat.element_size()

