"""Developer script (GPU box): random sample-rate pairs through the device resampler against the oracle (resampy restatement)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import ingest as oi  # noqa: E402

PAIRS = [(44100, 8000), (48000, 44100), (11025, 44100), (96000, 8000), (44100, 12345), (22050, 16000), (16000, 48000), (44100, 11025),
         (32000, 22050), (8000, 11025), (48000, 8000), (44100, 44100 // 3), (22050, 22051), (37800, 16000), (44100, 32000), (12000, 16000)]
bad = 0
for a, b in PAIRS:
    for f in ('kaiser_best', 'kaiser_fast'):
        n = int(a * 0.6) + 137
        x = piano_like(n, a, seed=a % 97)
        clips = [x, x[:2501], x[:11]]
        outs = ab.resample(clips, a, b, f)
        worst = 0.0
        for o, c in zip(outs, clips):
            want = oi.resample(c.astype(np.float64), a, b, f)
            got = o.cpu().numpy().astype(np.float64)
            if got.shape != want.shape:
                worst = 1.0
                continue
            if want.size:
                worst = max(worst, np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
        ok = worst <= 3e-7
        bad += not ok
        print('%s %d -> %d %s  max|d|/peak %.2e' % ('ok ' if ok else 'BAD', a, b, f, worst), flush=True)
print('misses:', bad)
sys.exit(1 if bad else 0)
