"""CPU suite: the multi-GPU plan is per-image sharding with no data-path collective.  Two gloo
ranks each take their shard_range of a job, "encode" it with the CPU oracle standing in for the GPU (test
infrastructure only), and the union must be every frame exactly once with the bytes a single rank produces.
Also covers the max-over-ranks reduction bench.py uses for its timing."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_range_partitions():
    import h2j_b200

    for n in (0, 1, 2, 7, 256, 100000):
        for world in (1, 2, 3, 4, 8):
            got = [h2j_b200.shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == n
            for a, b in zip(got, got[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in got]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        h2j_b200.shard_range(4, 2, 2)
    assert h2j_b200.sub_batches(3, 11, 4) == [(3, 7), (7, 11)]
    assert h2j_b200.sub_batches(5, 5, 4) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, w, h, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hashlib

    import h2j_b200
    from tests.support import oracle as orc

    lo, hi = h2j_b200.shard_range(n_frames, rank, world)
    digests = {}
    for a, b in h2j_b200.sub_batches(lo, hi, 2):
        for i in range(a, b):
            y, u, v = orc.synth_planes(w, h, "textured", seed=i, amp=20 + (i * 7) % 40)
            j, _, _ = orc.oracle_encode(y, u, v)
            digests[i] = hashlib.sha256(j).hexdigest()
    # what bench.py does with its timings: frames summed over ranks, time = max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([hi - lo])
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dist.barrier()
    q.put((rank, lo, hi, digests, float(t.item()), int(cnt.item())))
    dist.destroy_process_group()


def test_two_ranks_cover_the_job_once():
    from tests.support import oracle as orc
    import hashlib

    orc.oracle()
    n, w, h, world = 7, 48, 32, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, w, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = {}
    for rank, lo, hi, digests, tmax, cnt in res:
        assert tmax == float(world) and cnt == n
        assert sorted(digests) == list(range(lo, hi))
        for i, d in digests.items():
            assert i not in seen
            seen[i] = d
    assert sorted(seen) == list(range(n))
    for i in range(n):
        y, u, v = orc.synth_planes(w, h, "textured", seed=i, amp=20 + (i * 7) % 40)
        j, _, _ = orc.oracle_encode(y, u, v)
        assert seen[i] == hashlib.sha256(j).hexdigest()
