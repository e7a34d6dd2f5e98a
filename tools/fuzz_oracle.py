"""Developer script (GPU box): random CQT / VQT / HCQT / HVQT / STFT / MelSpec configurations, CUDA path against the float64 oracle
on ragged batches, at the bars of tests/test_gpu_parity.py.  Prints one line per configuration; exits non-zero on a miss.
usage: python tools/fuzz_oracle.py [nconf] [seed]"""
import os
import sys

import numpy as np

os.environ['AMTFEAT_DESCRIBE_ROWS'] = '1'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import modules as om  # noqa: E402
import rows_vs_oracle as rv  # noqa: E402


def tie_rows(m, kind, kw):
    """Rows whose sparsified kept set differs between the plan and the oracle: the 1 % threshold of librosa's sparsify_rows is decided
    by the float32 noise of the reference's own FFT there (Known deviations) -- no restatement can pin them."""
    if kind not in ('CQT', 'VQT', 'HCQT', 'HVQT'):
        return 0
    d = m.describe()
    harm = kw.get('harmonics', [1.0])
    fmin = kw.get('fmin') or rv.ls.NOTE_C1_HZ
    gamma = float(getattr(m, 'gamma', 0.0) or 0.0)
    n = 0
    for h, hv in enumerate(harm):
        want = rv.oracle_rows(kw['sample_rate'], kw['n_bins'], kw['bins_per_octave'], fmin * hv, gamma, d['eds_lib'][h])
        n += sum(1 for chan, b, col0, cnt, nnz in d['rows'] if chan == h and want[b] != (col0, cnt, nnz))
    return n


def rand_config(rng):
    kind = str(rng.choice(['CQT', 'VQT', 'HCQT', 'HVQT', 'STFT', 'MelSpec'], p=[0.2, 0.2, 0.2, 0.2, 0.1, 0.1]))
    sr = int(rng.choice([16000, 22050, 32000, 44100]))
    if kind in ('STFT', 'MelSpec'):
        n_fft = int(2 ** rng.randint(6, 12))
        hop = int(rng.choice([n_fft // 8, n_fft // 4, n_fft // 2, n_fft, 100, 160, 441]))
        kw = dict(sample_rate=sr, hop_length=max(1, hop), n_fft=n_fft, center=bool(rng.randint(0, 4) > 0))
        if rng.randint(0, 3) == 0:
            kw['win_length'] = int(rng.randint(n_fft // 4, n_fft + 1))
        if kind == 'MelSpec':
            kw['n_mels'] = int(rng.choice([40, 80, 128, 229]))
            kw['htk'] = bool(rng.randint(0, 2))
        return kind, kw
    bpo = int(rng.choice([12, 24, 36, 48, 60]))
    n_oct = int(rng.randint(2, 9))
    n_bins = bpo * n_oct - int(rng.randint(0, bpo // 2))
    hop = int(2 ** (n_oct - 1) * rng.choice([1, 2, 3, 4, 6, 8, 16]))
    while hop > 2048:
        hop //= 2
    fmin = float(rng.choice([27.5, 32.70319566257483, 41.2, 55.0, 65.4]))
    kw = dict(sample_rate=sr, hop_length=hop, n_bins=n_bins, bins_per_octave=bpo, fmin=fmin)
    if kind in ('VQT', 'HVQT'):
        g = rng.randint(0, 3)
        kw['gamma'] = None if g == 0 else float(rng.choice([0.0, 3.0, 11.0, 25.0]))
    if kind in ('HCQT', 'HVQT'):
        pool = [0.5, 1, 2, 3, 4, 5]
        k = int(rng.randint(2, 5))
        kw['harmonics'] = sorted(float(h) for h in rng.choice(pool, size=k, replace=False))
    return kind, kw


def main(nconf=30, seed=0):
    rng = np.random.RandomState(seed)
    bad = done = tried = 0
    while done < nconf and tried < 40 * nconf:
        tried += 1
        kind, kw = rand_config(rng)
        decibels = bool(rng.randint(0, 2))
        mk = lambda mod, extra: getattr(mod, ('O' if mod is om else '') + kind)(decibels=decibels, **dict({k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()}, **extra))
        try:
            m = mk(ab, {})
            m._dev_plan
        except Exception:          # invalid configuration (cutoff above Nyquist, n_fft out of range, ...)
            continue
        try:
            o, o32 = mk(om, {}), mk(om, dict(dtype=np.float32))
        except Exception as e:
            print('SKIP (oracle rejects what the plan accepts?)', kind, kw, repr(e)[:120], flush=True)
            continue
        sr = kw['sample_rate']
        n0 = int(rng.randint(sr // 3, int(sr * 2.5)))
        y = piano_like(n0, sr, seed=int(rng.randint(1 << 20)))
        clips = [y, y[:int(rng.randint(max(2, n0 // 8), n0))]]
        if not kw.get('center', True):
            need = kw.get('win_length') or kw.get('n_fft', 0)
            clips = [c for c in clips if len(c) >= need] or [y]
        got = m.process_audio(clips)
        worst_lin = worst_top = worst_all = worst_f32 = worst_f32_all = 0.0
        shape_ok = True
        for g, c in zip(got, clips):
            g = g.cpu().numpy().astype(np.float64)
            w = np.asarray(o.process_audio(c), np.float64)
            if g.shape != w.shape:
                shape_ok = False
                continue
            if not w.size:
                continue
            if decibels:
                d = np.abs(g - w) * 80.0
                top = w > 0.25
                worst_all = max(worst_all, d.max())
                w32 = np.asarray(o32.process_audio(c), np.float64)
                worst_f32_all = max(worst_f32_all, (np.abs(w32 - w) * 80.0).max())
                if top.any():
                    worst_top = max(worst_top, d[top].max())
                    worst_f32 = max(worst_f32, (np.abs(w32 - w) * 80.0)[top].max())
            else:
                worst_lin = max(worst_lin, np.linalg.norm(g - w) / max(np.linalg.norm(w), 1e-30))
        ok = shape_ok and worst_lin <= 1e-5 and worst_top <= max(1e-3, 1.5 * worst_f32) and worst_all <= max(2e-2, 1.5 * worst_f32_all)
        ties = 0 if ok else tie_rows(m, kind, kw)
        bad += (not ok) and ties == 0
        done += 1
        print('%s %-7s dB=%d %s lens=%s  lin=%.2e top=%.2e (f32 oracle %.2e) all=%.2e' % (
            'ok ' if ok else ('TIE (%d rows: sparsify threshold inside float32 FFT noise)' % ties if ties else 'BAD'), kind, decibels, {k: v for k, v in kw.items()}, [len(c) for c in clips], worst_lin, worst_top, worst_f32, worst_all), flush=True)
    print('configurations: %d, misses: %d' % (done, bad))
    return bad


if __name__ == '__main__':
    sys.exit(1 if main(*(int(a) for a in sys.argv[1:])) else 0)
