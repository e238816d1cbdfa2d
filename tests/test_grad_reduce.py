"""Row a13 (train_gpt.py:672,798 DDP gradient all-reduce): the bucketed, overlapped reducer on CPU with gloo, world 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _grads(rank):
    g = torch.Generator().manual_seed(100 + rank)
    big = torch.randn(6, 8, generator=g)
    return {"lm_head.weight": torch.randn(5, 4, generator=g), "model.norm.weight": torch.randn(4, generator=g),
            "model.layers.0.mlp.gate_proj.weight": big[0::2], "model.layers.0.mlp.up_proj.weight": big[1::2],   # strided views
            "model.embed_tokens.weight": torch.randn(7, 4, generator=g)}


ORDER = [["lm_head.weight", "model.norm.weight"],
         ["model.layers.0.mlp.gate_proj.weight", "model.layers.0.mlp.up_proj.weight"], ["model.embed_tokens.weight"]]


def _worker(rank, world, port, out, min_bytes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ivideogpt_b200.grad_reduce import BucketedGradReducer
    red = BucketedGradReducer(min_bucket_bytes=min_bytes)
    grads = _grads(rank)
    for names in ORDER:
        red.on_grads(grads, names)
    red.finish(grads)
    # numpy, not tensors: a tensor travels through the queue as a shared-memory handle served by THIS process, which may
    # have exited before the parent opens it
    out.put((rank, {k: v.numpy().copy() for k, v in grads.items()}, red.buckets_launched))
    dist.destroy_process_group()


def _run(min_bytes):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, min_bytes)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_bucketed_reducer_equals_plain_sum_gloo():
    want = {k: _grads(0)[k] + _grads(1)[k] for k in _grads(0)}
    for min_bytes, nb in ((0, 3), (10 ** 9, 1)):         # one bucket per call / everything merged into one bucket
        res = _run(min_bytes)
        for rank, got, buckets in res:
            assert buckets == nb
            assert set(got) == set(want)
            for k in want:
                assert tuple(got[k].shape) == tuple(want[k].shape) and torch.equal(torch.from_numpy(got[k]), want[k]), (rank, k)


def test_reducer_single_process_is_identity():
    from ivideogpt_b200.grad_reduce import BucketedGradReducer
    red = BucketedGradReducer()
    grads = _grads(0)
    want = {k: v.clone() for k, v in grads.items()}
    for names in ORDER:
        red.on_grads(grads, names)
    red.finish(grads)
    assert all(torch.equal(grads[k], want[k]) and grads[k].is_contiguous() for k in want)


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ivideogpt_b200.grad_reduce import allreduce_grads_flat
    g = torch.Generator().manual_seed(7 + rank)
    params = [torch.nn.Parameter(torch.zeros(3, 5)), torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 2, 3))]
    params[0].grad = torch.randn(3, 5, generator=g)
    params[2].grad = torch.randn(2, 3, 2, generator=g).permute(0, 2, 1)      # a non-contiguous gradient; params[1] has none
    nbytes = allreduce_grads_flat(params)
    out.put((rank, [None if p.grad is None else p.grad.contiguous().numpy().copy() for p in params], nbytes))
    dist.destroy_process_group()


def test_flat_gradient_exchange_gloo():
    """The tokenizer training step's exchange (bench.py train_tokenizer64 leg): one flat SUM over the existing gradients."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    def local(rank):
        g = torch.Generator().manual_seed(7 + rank)
        return torch.randn(3, 5, generator=g), torch.randn(2, 3, 2, generator=g).permute(0, 2, 1)
    w0, w2 = local(0)[0] + local(1)[0], local(0)[1] + local(1)[1]
    for rank, grads, nbytes in res:
        assert nbytes == (15 + 12) * 4 and grads[1] is None
        assert torch.equal(torch.from_numpy(grads[0]), w0) and torch.equal(torch.from_numpy(grads[2]), w2.contiguous())
    from ivideogpt_b200.grad_reduce import allreduce_grads_flat
    p = torch.nn.Parameter(torch.zeros(2)); p.grad = torch.ones(2)
    assert allreduce_grads_flat([p]) == 0 and torch.equal(p.grad, torch.ones(2))        # single process: untouched
