import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
from oracle import modules as om

def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)

y = piano_like(22050 * 5, 22050, seed=211)
kw = dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24, decibels=False)
fast = ab.CQT(**kw).process_audio(y).cpu().numpy().astype(np.float64)
os.environ['AMTFEAT_DECIM'] = 'direct'
direct = ab.CQT(**kw).process_audio(y).cpu().numpy().astype(np.float64)
want = om.OCQT(**kw).process_audio(y)
print('fast vs oracle', rel(fast, want), 'direct vs oracle', rel(direct, want), 'fast vs direct', rel(fast, direct))
for o in range(8):
    sl = slice(24 * o, 24 * o + 24)
    d = np.abs(fast[0, sl] - direct[0, sl])
    t = d.max(axis=0)
    print('octave', o, 'fast/oracle %.3g direct/oracle %.3g fast/direct %.3g' % (rel(fast[0, sl], want[0, sl]), rel(direct[0, sl], want[0, sl]), rel(fast[0, sl], direct[0, sl])),
          'worst frames', np.argsort(-t)[:6], 'T', fast.shape[-1])
