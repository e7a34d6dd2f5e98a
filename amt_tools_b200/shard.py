"""
Track-level sharding of a corpus over the GPUs of one box.  `process_audio` has no
cross-track state (features/common.py:168-179: every override is a pure function of the audio and the
constructor arguments), so the unit of partitioning is the track and the data path needs no collective;
only the throughput figure is reduced (a sum of audio seconds and a max of elapsed times).
"""

import torch
import torch.distributed as dist


def shard_tracks(lengths, world_size, rank):
    """Indices of the tracks rank `rank` processes: longest-first greedy assignment to the least-loaded rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        loads[r] += int(lengths[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def make_batches(indices, lengths, max_samples):
    """Groups a rank's tracks into ragged batches of at most `max_samples` samples (a lone longer track is its own batch)."""
    batches, cur, cur_n = [], [], 0
    for i in indices:
        n = int(lengths[i])
        if cur and cur_n + n > max_samples:
            batches.append(cur)
            cur, cur_n = [], 0
        cur.append(i)
        cur_n += n
    if cur:
        batches.append(cur)
    return batches


def aggregate_throughput(audio_seconds, elapsed_seconds):
    """(sum over ranks of audio seconds, max over ranks of elapsed seconds); identity without a process group."""
    total = audio_seconds.clone()
    slowest = elapsed_seconds.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
        dist.all_reduce(slowest, op=dist.ReduceOp.MAX)
    return total, slowest


def bind_host_to_gpu(device_index):
    """
    Pins the calling process to the CPU cores next to GPU `device_index` (NVML's ideal affinity), so the pinned staging
    buffers a rank allocates afterwards land on the NUMA node its PCIe root port hangs off: with one rank per GPU the
    device-to-host copies of the features (5x the audio) would otherwise all target the node the launcher started on.
    Returns the affinity list, or None when NVML / the cpuset does not allow it (the data path is unaffected either way).
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
            cpus = [64 * w + b for w, mask in enumerate(words) for b in range(64) if (int(mask) >> b) & 1]
            allowed = os.sched_getaffinity(0)
            cpus = sorted(c for c in cpus if c in allowed)
            if not cpus:
                return None
            os.sched_setaffinity(0, cpus)
            return cpus
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return None
