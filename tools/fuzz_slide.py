"""Developer script (GPU box): random CQT / VQT / HVQT configurations, sliding-DFT kernel against the FFT-per-frame kernel
(AMTFEAT_SLIDE=0) on ragged batches.  Prints one line per configuration; exits non-zero on a mismatch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402


def main(nconf=24, seed=0):
    rng = np.random.RandomState(seed)
    bad = 0
    done = 0
    while done < nconf:
        sr = int(rng.choice([16000, 22050, 44100]))
        bpo = int(rng.choice([12, 24, 36, 60]))
        n_oct = int(rng.randint(3, 9))
        n_bins = bpo * n_oct - int(rng.randint(0, bpo // 2))
        hop = int(2 ** rng.randint(5, 11))
        kind = rng.choice(['CQT', 'VQT', 'HVQT'])
        fmin = float(rng.choice([27.5, 32.70319566257483, 55.0, 65.4]))
        kw = dict(sample_rate=sr, hop_length=hop, n_bins=n_bins, bins_per_octave=bpo, fmin=fmin)
        if kind == 'VQT':
            kw['gamma'] = float(rng.choice([0.0, 5.0, 20.0]))
        if kind == 'HVQT':
            kw['harmonics'] = [1, 2, 3][:int(rng.randint(2, 4))]
            kw['gamma'] = 0.0
        try:
            os.environ.pop('AMTFEAT_SLIDE', None)
            slide = getattr(ab, kind)(decibels=False, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
            items = slide.describe()['items']
            slide._dev_plan     # the switch is read at plan creation; the device plan is created lazily
        except Exception as e:     # invalid configuration (cutoff above Nyquist, hop not divisible, ...): same on both paths
            continue
        if not any(it['slide'] for it in items):
            continue
        os.environ['AMTFEAT_SLIDE'] = '0'
        plain = getattr(ab, kind)(decibels=False, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
        plain.describe()
        plain._dev_plan
        os.environ.pop('AMTFEAT_SLIDE', None)
        n0 = int(rng.randint(sr // 2, sr * 6))
        y = piano_like(n0, sr, seed=int(rng.randint(1 << 20)))
        clips = [y, y[:int(rng.randint(1, n0))], y[:int(rng.randint(1, 4 * hop))]]
        a, b = slide.process_audio(clips), plain.process_audio(clips)
        worst = 0.0
        for u, v in zip(a, b):
            assert u.shape == v.shape
            if u.numel():
                u, v = u.double().cpu().numpy(), v.double().cpu().numpy()
                worst = max(worst, float(np.abs(u - v).max() / max(np.abs(v).max(), 1e-30)))
        ok = 0.0 < worst < 5e-6          # 0 would mean both plans took the same path
        bad += not ok
        done += 1
        print('%s %-5s sr=%d hop=%d bins=%d bpo=%d fmin=%.1f %s slide-items=%d/%d (hops %s)  max|d|/peak=%.2e' % (
            'ok ' if ok else 'BAD', kind, sr, hop, n_bins, bpo, fmin, {k: kw[k] for k in ('gamma', 'harmonics') if k in kw},
            sum(it['slide'] for it in items), len(items), sorted({it['hop'] for it in items if it['slide']}), worst), flush=True)
    return bad


if __name__ == '__main__':
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:])) else 0)
