"""Small ragged batches through every kernel family, for compute-sanitizer (memcheck / racecheck) on the GPU box."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like

y = piano_like(22050 * 3 + 77, 22050, seed=1)
mods = (ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60), ab.CQT(22050, 128, n_bins=96, bins_per_octave=12), ab.MelSpec(), ab.VQT(22050, 512),
        ab.HCQT(22050, 512, harmonics=[0.25, 0.5, 1], n_bins=96, bins_per_octave=24))
for m in mods:
    out = m.process_audio([y, y[:20011], y[:700], y[:23807]])
    torch.cuda.synchronize()
    print(type(m).__name__, [tuple(o.shape) for o in out], float(out[0].max()))
