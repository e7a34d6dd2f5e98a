"""Developer script (GPU box): plans, resamplers and pipelines created and destroyed in a loop must give their device memory back."""
import gc
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200 import ingest
from amt_tools_b200.pipeline import Pipeline
from amt_tools_b200.synth import piano_like

y = piano_like(22050 * 2, 22050, seed=1)
def used():
    torch.cuda.synchronize()
    gc.collect()
    torch.cuda.empty_cache()
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 1e6
for _ in range(3):      # warm: context, caches
    m = ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60); m.process_audio(y); del m
    r = ingest.Resampler(44100, 16000); r(y); del r
base = used()
for i in range(60):
    m = ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60); m.process_audio(y); del m
    m = ab.MelSpec(16000); m.process_audio(y); del m
    r = ingest.Resampler(44100, 16000); r(y); del r
    p = Pipeline(0, 2, 1 << 20, 1 << 22, 1 << 24); del p
after = used()
print('device memory in use: %.1f MB before, %.1f MB after 60 create / destroy rounds (delta %.1f MB)' % (base, after, after - base))
sys.exit(1 if after - base > 64 else 0)
