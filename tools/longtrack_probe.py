"""Developer script (GPU box): ONE long track through amt_tools_b200.longtrack on 1 .. N GPUs (python tools/longtrack_probe.py, or under
torchrun --nproc-per-node N): whole-track process_audio on one GPU against the chunked path (chunks dealt to the ranks, one all_reduce(MAX)
of C floats, optional gather), audio resident on every rank's device.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200 import longtrack as lt
from amt_tools_b200.synth import piano_like

minutes = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
dev = torch.device('cuda', local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
group = dist.group.WORLD if world > 1 else None

sr = 22050
n = int(sr * 60 * minutes)
seg = piano_like(sr * 60, sr, seed=2)
y = np.tile(seg, n // len(seg) + 1)[:n].copy()
y *= (1.0 + 0.1 * np.sin(np.arange(n, dtype=np.float32) * 1e-6)).astype(np.float32)
yd = torch.from_numpy(y).to(dev)
res = {'track_minutes': minutes, 'n_gpus': world}
for name, kw in (('HCQT', dict(sample_rate=sr, hop_length=256, n_bins=360, bins_per_octave=60)), ('MelSpec', dict(sample_rate=sr, hop_length=512, n_fft=2048))):
    m = getattr(ab, name)(device=dev, **kw)

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, out

    t_whole, whole = timed(lambda: m.process_audio(yd))
    cf = max(16 * lt.ALIGN, -(-int(m._out_shape(n)[-1]) // (4 * world)))
    t_own, own = timed(lambda: lt.process_long_audio(m, yd, chunk_frames=cf, group=group, gather=False))
    # the rank's own chunks against the whole-track result
    worst = 0.0
    for ci, (f0, f1, part) in own.items():
        w = whole[..., f0:f1]
        sel = w > 0.25
        worst = max(worst, float((part - w).abs()[sel].max()) * 80.0)
    del own
    t_all, full = timed(lambda: lt.process_long_audio(m, yd, chunk_frames=cf, group=group, gather=True), reps=2)
    same_shape = tuple(full.shape) == tuple(whole.shape)
    del full, whole
    torch.cuda.empty_cache()
    w = torch.tensor([worst], device=dev)
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    hours = minutes / 60.0
    res[name] = {'whole_track_1gpu_s': round(t_whole, 4), 'chunked_s': round(t_own, 4), 'chunked_gathered_s': round(t_all, 4),
                 'whole_audio_h_per_s': round(hours / t_whole, 1), 'chunked_audio_h_per_s': round(hours / t_own, 1),
                 'speedup_vs_whole': round(t_whole / t_own, 2), 'max_db_diff_graded_bins': float(w), 'shape_ok': same_shape,
                 'halo_frames': lt.halo_frames(m, 4096)}
if rank == 0:
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
