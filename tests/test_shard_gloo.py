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


def test_weighted_shards_partition_in_proportion():
    import h2j_b200

    # the pool's 8-GPU box: four links at 23.6 GB/s, four at 36.1; 8 x 2048 frames in sub-batches of 64
    rates = [23.6] * 4 + [36.1] * 4
    sh = h2j_b200.weighted_shards(8 * 2048, rates, 64)
    sizes = [b - a for a, b in sh]
    assert sh[0][0] == 0 and sh[-1][1] == 8 * 2048 and all(a[1] == b[0] for a, b in zip(sh, sh[1:]))
    assert all(s % 64 == 0 for s in sizes) and sum(sizes) == 8 * 2048
    assert sizes[:4] == [1600] * 4 or abs(sizes[0] - 1619) <= 64
    assert all(abs(s / 64 - 256 * r / sum(rates)) <= 1 for s, r in zip(sizes, rates))  # within one sub-batch of the ideal share
    # equal weights: shard_range's split
    assert h2j_b200.weighted_shards(4 * 2048, [55.0] * 4, 64) == [h2j_b200.shard_range(4 * 2048, r, 4) for r in range(4)]
    # ragged totals: the remainder rides with the last rank; nobody is left empty while there is work for everybody
    for n in (0, 1, 5, 63, 64, 65, 1000, 4097):
        for w in ([1.0], [1.0, 3.0], [5.0, 1.0, 1.0], [0.1, 10.0, 10.0, 10.0]):
            for g in (1, 2, 64):
                sh = h2j_b200.weighted_shards(n, w, g)
                assert sh[0][0] == 0 and sh[-1][1] == n and all(a[1] == b[0] for a, b in zip(sh, sh[1:])) and all(a <= b for a, b in sh)
                assert all((b - a) % g == 0 for a, b in sh[:-1])
                if n // g >= len(w):
                    assert all(b > a for a, b in sh)
    with pytest.raises(ValueError):
        h2j_b200.weighted_shards(10, [1.0, 0.0])
    with pytest.raises(ValueError):
        h2j_b200.weighted_shards(10, [])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, w, h, q, weighted):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import hashlib

    import h2j_b200
    from tests.support import oracle as orc

    # rank 0 sits behind the slower link in this story: the weighted split gives it the smaller share
    lo, hi = h2j_b200.weighted_shards(n_frames, [1.0, 2.5], 1)[rank] if weighted else h2j_b200.shard_range(n_frames, rank, world)
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


@pytest.mark.parametrize("weighted", [False, True])
def test_two_ranks_cover_the_job_once(weighted):
    from tests.support import oracle as orc
    import hashlib

    orc.oracle()
    n, w, h, world = 7, 48, 32, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, w, h, q, weighted)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seen = {}
    if weighted:
        assert sorted((hi - lo) for _, lo, hi, _, _, _ in res) == [2, 5]
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
