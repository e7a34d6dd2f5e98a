"""Developer probe: throughput of the device audio ingest (resample + rms_norm) on 240 s tracks."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from amt_tools_b200 import ingest
from amt_tools_b200.synth import piano_like
dev = torch.device('cuda', 0)
B = 8
for src, dst in ((44100, 22050), (44100, 16000), (48000, 22050)):
    y = [torch.from_numpy(piano_like(src * 240, src, seed=i % 2)).to(dev) for i in range(B)]
    rs = ingest.Resampler(src, dst, device=dev)
    for _ in range(2):
        o = rs(y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        o = rs(y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    e0.record()
    for _ in range(5):
        z = [ingest.rms_norm(q, device=dev) for q in o]
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / 5
    print('%d -> %d: resample %.3f ms per %d x 240 s (%.0f audio-h/s), rms_norm %.3f ms' % (src, dst, ms, B, B * 240 / 3600 / (ms * 1e-3), ms2))
