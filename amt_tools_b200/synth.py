"""
Seeded synthetic "piano-like" audio used by the tests and the benchmark:
~2 notes/s, f0 = 27.5 * 2^(k/12) with k ~ U{20..80}, 7 harmonics at 1/h amplitude,
exponential decay tau ~ U(0.1, 0.7) s, plus white noise at -60 dBFS, then RMS-normalised
the way the reference normalises loaded audio (amt_tools/tools/utils.py:2806-2812).
"""

import numpy as np


def piano_like(num_samples, sample_rate, seed=0, notes_per_second=2.0):
    rng = np.random.RandomState(seed)
    dur = num_samples / float(sample_rate)
    n_notes = max(1, int(round(dur * notes_per_second)))
    y = np.zeros(num_samples, dtype=np.float64)
    onsets = np.sort(rng.uniform(0, max(dur - 0.05, 0.0), n_notes))
    keys = rng.randint(20, 81, n_notes)
    taus = rng.uniform(0.1, 0.7, n_notes)
    amps = rng.uniform(0.2, 1.0, n_notes)
    for onset, k, tau, a in zip(onsets, keys, taus, amps):
        i0 = int(onset * sample_rate)
        i1 = min(num_samples, i0 + int(6 * tau * sample_rate))
        if i1 <= i0:
            continue
        t = np.arange(i1 - i0) / float(sample_rate)
        f0 = 27.5 * 2.0 ** (k / 12.0)
        env = a * np.exp(-t / tau)
        note = np.zeros_like(t)
        for h in range(1, 8):
            if h * f0 < 0.45 * sample_rate:
                note += np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 2 * np.pi)) / h
        y[i0:i1] += env * note
    y += 1e-3 * rng.randn(num_samples)
    rms = np.sqrt(np.mean(y ** 2))
    if rms > 0:
        y = y / rms
    return y.astype(np.float32)
