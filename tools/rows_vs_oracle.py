"""Host-only check (no GPU): the sparsified wavelet rows of the C++ plan (first kept FFT bin, band width, kept entries) against the
oracle's `vqt_filter_fft` for random CQT / VQT configurations.  usage: AMTFEAT_DESCRIBE_ROWS=1 python tools/rows_vs_oracle.py [nconf] [seed]"""
import os
import sys

import numpy as np

os.environ['AMTFEAT_DESCRIBE_ROWS'] = '1'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amt_tools_b200 as ab  # noqa: E402
from oracle import librosa_stages as ls  # noqa: E402


def oracle_rows(sr, n_bins, bpo, fmin, gamma, eds):
    """(bin, col0, cnt, nnz) of every row, following the octave recursion of librosa.vqt (oracle/librosa_stages.py)."""
    freqs = ls.cqt_frequencies(n_bins, fmin, bpo)
    alpha = ls.relative_bandwidth_et(bpo)
    n_oct = int(np.ceil(n_bins / bpo))
    n_filters = min(bpo, n_bins)
    out = {}
    for i in range(n_oct):
        lo, hi = max(0, n_bins - n_filters * (i + 1)), n_bins - n_filters * i
        my_sr = sr / 2.0 ** (eds + i)
        fb, n_fft, _ = ls.vqt_filter_fft(my_sr, freqs[lo:hi], gamma, alpha)
        fb = fb.toarray()
        for r in range(hi - lo):
            nz = np.nonzero(fb[r])[0]
            out[lo + r] = (int(nz.min()), int(nz.max() - nz.min() + 1), int(len(nz)))
    return out


def main(nconf=40, seed=0):
    rng = np.random.RandomState(seed)
    done = mism = rows_total = 0
    while done < nconf:
        sr = int(rng.choice([16000, 22050, 32000, 44100]))
        bpo = int(rng.choice([12, 24, 36, 48, 60]))
        n_oct = int(rng.randint(2, 9))
        n_bins = bpo * n_oct - int(rng.randint(0, bpo // 2))
        hop = int(2 ** (n_oct - 1) * rng.choice([1, 2, 4, 8]))
        fmin = float(rng.choice([27.5, 32.70319566257483, 41.2, 55.0, 65.4, 36.7]))
        gamma = float(rng.choice([0.0, 0.0, 3.0, 11.0, 25.0]))
        try:
            m = ab.VQT(sample_rate=sr, hop_length=hop, n_bins=n_bins, bins_per_octave=bpo, fmin=fmin, gamma=gamma)
            d = m.describe()
        except Exception:
            continue
        eds = d['eds_lib'][0]
        want = oracle_rows(sr, n_bins, bpo, fmin, gamma, eds)
        bad = []
        for chan, b, col0, cnt, nnz in d['rows']:
            if want[b] != (col0, cnt, nnz):
                bad.append((b, (col0, cnt, nnz), want[b]))
        rows_total += len(d['rows'])
        mism += len(bad)
        done += 1
        print('%s sr=%d bins=%d bpo=%d fmin=%.2f gamma=%.0f hop=%d eds=%d rows=%d mismatches=%d %s' % (
            'ok ' if not bad else 'DIFF', sr, n_bins, bpo, fmin, gamma, hop, eds, len(d['rows']), len(bad), bad[:3]), flush=True)
    print('rows %d, mismatching kept sets %d' % (rows_total, mism))
    return mism


if __name__ == '__main__':
    main(*(int(a) for a in sys.argv[1:]))
