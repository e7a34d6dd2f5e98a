"""One step of a bench workload between cudaProfilerStart / Stop, for `ncu --profile-from-start off --set full ...`.
usage: python tools/one_step.py c5|c2|c3|c4"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import amt_tools_b200 as ab
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else 'c5'
desc, spec, seconds, B = bench.WORKLOADS[wl]
dev = torch.device('cuda', 0)
mods = [getattr(ab, name)(device=dev, **kw) for name, kw, sr in spec]
audio = [torch.from_numpy(bench.synth_batch(sr, seconds, B, seed0=100 + i)).to(dev) for i, (name, kw, sr) in enumerate(spec)]
for _ in range(2):
    outs = [m.process_audio(a) for m, a in zip(mods, audio)]
torch.cuda.synchronize()
torch.cuda.profiler.start()
outs = [m.process_audio(a) for m, a in zip(mods, audio)]
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(wl, desc, [tuple(o.shape) for o in outs])
