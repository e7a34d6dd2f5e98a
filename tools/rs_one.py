import sys, numpy as np, torch
sys.path.insert(0, '.')
from amt_tools_b200 import ingest
from amt_tools_b200.synth import piano_like
dev = torch.device('cuda', 0)
y = [torch.from_numpy(piano_like(44100 * 240, 44100, seed=i % 2)).to(dev) for i in range(8)]
rs = ingest.Resampler(44100, 16000, device=dev)
o = rs(y); o = rs(y)
torch.cuda.synchronize()
