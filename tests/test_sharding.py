"""
Track sharding across ranks: no collective on the data path; the union of the per-rank
shards is the whole corpus, balanced longest-first.  The N > 1 plumbing is covered with world_size-2 gloo on CPU.
"""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from amt_tools_b200 import shard


def test_shard_tracks_partition_and_balance():
    rng = np.random.RandomState(0)
    lengths = rng.randint(22050 * 30, 22050 * 600, size=101).tolist()
    for world in (1, 2, 4, 8):
        parts = [shard.shard_tracks(lengths, world, r) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))                       # a partition: every track exactly once
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)                  # longest-first greedy bound
    assert shard.shard_tracks([], 4, 2) == []
    assert shard.shard_tracks([5, 5, 5], 8, 7) == []


def test_batches_respect_budget_and_order():
    lengths = [100, 900, 400, 400, 50, 1000]
    idx = list(range(len(lengths)))
    batches = shard.make_batches(idx, lengths, max_samples=1000)
    assert sorted(i for b in batches for i in b) == idx
    assert all(sum(lengths[i] for i in b) <= 1000 or len(b) == 1 for b in batches)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, lengths, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mine = shard.shard_tracks(lengths, world, rank)
    seconds = torch.tensor([float(sum(lengths[i] for i in mine)) / 22050.0], dtype=torch.float64)
    elapsed = torch.tensor([0.5 + 0.25 * rank], dtype=torch.float64)           # stand-in for the device-timed region
    total, slowest = shard.aggregate_throughput(seconds, elapsed)
    counts = [None] * world
    dist.all_gather_object(counts, mine)
    if rank == 0:
        out.put((float(total), float(slowest), counts))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    lengths = [22050 * s for s in (240, 200, 180, 120, 90, 60, 30)]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    total, slowest, counts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(i for c in counts for i in c) == list(range(len(lengths)))
    assert abs(total - sum(lengths) / 22050.0) < 1e-9 and slowest == 0.75


def test_host_binding_is_optional():
    # no NVML / no GPU here: the helper must decline quietly and leave the affinity alone
    import os
    from amt_tools_b200 import shard
    before = os.sched_getaffinity(0)
    cpus = shard.bind_host_to_gpu(0)
    assert cpus is None or set(cpus) <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
