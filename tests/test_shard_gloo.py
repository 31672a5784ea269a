"""N > 1 path on CPU: world_size-2 gloo run of the pieces bench.py uses across ranks (partition by session id,
barrier, max-over-ranks of the step time, whole-job aggregation). No collective touches audio data."""
import os
import socket
import uuid

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streamkit_b200 import shard


def test_fnv1a_known_answers_and_partition():
    assert shard.fnv1a64(b"") == 0xCBF29CE484222325
    assert shard.fnv1a64(b"a") == 0xAF63DC4C8601EC8C
    ids = [str(uuid.UUID(int=i * 7919 + 13)) for i in range(20000)]
    for n in (1, 2, 4, 8):
        parts = shard.partition(ids, n)
        assert sum(len(p) for p in parts) == len(ids)
        assert sorted(x for p in parts for x in p) == sorted(ids)
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) < 0.1 * len(ids) / n + 50     # near-even split
        for r, p in enumerate(parts):
            assert all(shard.gpu_for_session(s, n) == r for s in p[:50])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = [str(uuid.UUID(int=i * 7919 + 13)) for i in range(4000)]
        mine = shard.partition(ids, world)[rank]
        # every rank "processes" its shard; the slowest rank defines the step time (bench.py: max over ranks)
        my_ms = 1.0 + rank * 0.5
        dist.barrier()
        t = torch.tensor([my_ms], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n = torch.tensor([len(mine)], dtype=torch.int64)
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        if rank == 0:
            out.put((float(t.item()), int(n.item()), shard.aggregate(2000, world, float(t.item()))))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    max_ms, total, agg = q.get(timeout=90)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert max_ms == 1.5 and total == 4000
    assert agg == pytest.approx(2000 * 2 * 20.0 / 1.5)
