"""Developer probe: MelSpec 64 x 20 s (BASELINE configs[1]) device-resident step time; AMTFEAT_STFT_TPC=n forces the tiles per CTA."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
dev = torch.device('cuda', 0)
B = 64
m = ab.MelSpec(device=dev)
y = np.stack([piano_like(16000 * 20, 16000, seed=i) for i in range(4)] * (B // 4))
copies = [torch.from_numpy(y).to(dev) for _ in range(5)]
outs = []
for i in range(10):
    outs.append(m.process_audio(copies[i % 5]))
torch.cuda.synchronize()
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        outs[i % 10] = m.process_audio(copies[i % 5])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20)
print('c2 step %.4f ms  (%.0f audio-h/s)' % (best, B * 20 / 3600 / (best * 1e-3)))
