"""N > 1 plumbing on CPU: world_size-2 gloo run of the replica sharding + max-over-ranks timing protocol."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ivideogpt_b200 import replicas
    lo, hi = replicas.shard_range(13, rank, world)
    replicas.barrier()
    t = replicas.max_over_ranks(10.0 + 5.0 * rank)
    # every clip is owned exactly once: gather the ranges
    got = [None] * world
    dist.all_gather_object(got, (lo, hi))
    out.put((rank, (lo, hi), t, got))
    dist.destroy_process_group()


def test_two_rank_gloo_replicas():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert res[0][1] == (0, 7) and res[1][1] == (7, 13)
    assert res[0][2] == 15.0 and res[1][2] == 15.0            # MAX over ranks, identical on both
    assert res[0][3] == [(0, 7), (7, 13)]


def test_shard_range_properties():
    from ivideogpt_b200.replicas import shard_range
    for g in (0, 1, 7, 64, 256):
        for w in (1, 2, 3, 8):
            spans = [shard_range(g, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == g
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
