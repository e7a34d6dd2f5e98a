"""Developer script: one failing fuzz configuration in detail."""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
from oracle import modules as om

kw = dict(sample_rate=22050, hop_length=128, n_bins=299, bins_per_octave=60, fmin=55.0)
y = piano_like(25814, 22050, seed=6)
def run(H, tag):
    m = ab.HCQT(decibels=False, harmonics=list(H), **kw)
    o = om.OHCQT(decibels=False, harmonics=list(H), **kw)
    g = m.process_audio(y).cpu().numpy().astype(np.float64)
    w = np.asarray(o.process_audio(y), np.float64)
    for c in range(g.shape[0]):
        e = np.abs(g[c] - w[c]) / w[c].max()
        bad = np.argwhere(e > 2e-6)
        bins = sorted(set(bad[:, 0].tolist()))
        fr = sorted(set(bad[:, 1].tolist()))
        print(tag, 'H', H, 'chan', c, 'rel_l2 %.3g' % (np.linalg.norm(g[c] - w[c]) / np.linalg.norm(w[c])), 'bad cells', len(bad), 'bins', bins[:12], 'frames', fr[:6], '..', fr[-6:] if fr else '')
    return m
m = run([1.0, 2.0, 4.0, 5.0], 'default')
for it in m.describe()['items']:
    print(it)
run([1.0], 'default')
run([2.0], 'default')
run([1.0, 2.0], 'default')
