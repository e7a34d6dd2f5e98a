"""
Numerics of the sliding-DFT form of K1/K5 (cqt_slide_kernel), restated in float32 numpy: the frame
spectra it produces must stay at FFT-level accuracy against a float64 FFT of the same frames -- including right after a
loud passage has left the window (the sliding sum is only ever added to, Kahan-compensated; the phase is re-seeded from
the exact table every 8 frames; every tile starts from zero with a lead-in of one window).  CPU only; the CUDA kernel is
compared with the FFT-per-frame kernel and with the oracle in tests/test_gpu_parity.py.
"""
import numpy as np

f32, c64 = np.float32, np.complex64


def _signal(n, seed):
    rng = np.random.RandomState(seed)
    t = np.arange(n)
    y = np.zeros(n)
    for i in range(12):                      # decaying partials inside the band: loud ones early, quiet ones late
        on, tau = rng.randint(0, n // 3) + (n // 2 if i % 2 else 0), rng.uniform(0.01, 0.04) * n
        f, a = rng.uniform(0.09, 0.29), (10.0 ** -rng.uniform(2.0, 3.0) if i % 2 else rng.uniform(0.3, 1.0))
        y += np.where(t >= on, np.exp(-(t - on) / tau), 0) * a * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
    y += 1e-5 * rng.randn(n)
    return (y / np.sqrt(np.mean(y ** 2))).astype(f32)


def _frames(x, N, h, T, dtype):
    xp = np.concatenate([np.zeros(N // 2, dtype), x.astype(dtype), np.zeros(N + h * T, dtype)])
    return np.stack([xp[t * h: t * h + N] for t in range(T)])


def _slide(x, N, h, T, kmin, kmax, tile, reseed, kahan=True):
    k = np.arange(kmin, kmax + 1)
    tw = np.exp(-2j * np.pi * np.arange(N) / N).astype(c64)          # the kernel's table: float64 values rounded once
    Wm = [tw[(k * m) % N] for m in range(h)]
    R, Q = tw[(k * h) % N], N // h
    out = np.zeros((T, len(k)), c64)

    def get(j0):
        idx = np.arange(j0, j0 + h)
        ok = (idx >= 0) & (idx < len(x))
        v = np.zeros(h, f32)
        v[ok] = x[idx[ok]]
        return v

    for t0 in range(0, T, tile):
        B, C, P = np.zeros(len(k), c64), np.zeros(len(k), c64), None
        for u in range(Q + tile):
            t = t0 - Q + u
            if u % reseed == 0:
                P = tw[(k * ((t * h) % N)) % N]
            if u >= Q:                                               # main part: emit X_t, then enter / leave samples
                if t < T:
                    out[t] = (np.conj(P) * B).astype(c64)
                d = (get(t * h + N // 2) - get(t * h - N // 2)).astype(f32)
            else:                                                    # lead-in: the first window enters
                d = get(t * h + N // 2)
            delta = np.zeros(len(k), c64)
            for m in range(h):
                delta = (delta + d[m] * Wm[m]).astype(c64)
            term = (P * delta).astype(c64)
            if kahan:
                y = (term - C).astype(c64)
                nB = (B + y).astype(c64)
                C = ((nB - B).astype(c64) - y).astype(c64)
                B = nB
            else:
                B = (B + term).astype(c64)
            P = (P * R).astype(c64)
    return out


def _errors(got, want):
    mag, ref = np.abs(got.astype(np.complex128)), np.abs(want)
    rel = np.linalg.norm(mag - ref) / np.linalg.norm(ref)
    db_g, db_w = 20 * np.log10(np.maximum(mag, 1e-5)), 20 * np.log10(np.maximum(ref, 1e-5))
    top = db_w > db_w.max() - 60
    return rel, np.abs(db_g - db_w)[top].max()


def test_sliding_dft_keeps_fft_level_accuracy():
    import torch
    for N, h, kmin, kmax, tile in ((1024, 8, 94, 300, 512), (512, 16, 94, 248, 256), (1024, 2, 94, 200, 1024)):
        T = 1200
        x = _signal(T * h, seed=h)
        want = np.fft.rfft(_frames(x, N, h, T, np.float64), axis=1)[:, kmin:kmax + 1]
        dyn = 20 * np.log10(np.abs(want).max(1).max() / np.abs(want).max(1).min())
        assert dyn > 40                                              # loud and quiet frames inside the same tiles
        fft_rel, fft_db = _errors(torch.fft.rfft(torch.from_numpy(_frames(x, N, h, T, f32)), dim=1).numpy()[:, kmin:kmax + 1], want)
        rel8, db8 = _errors(_slide(x, N, h, T, kmin, kmax, tile, 8), want)
        rel32, _ = _errors(_slide(x, N, h, T, kmin, kmax, tile, 32), want)
        assert rel8 < 4e-7 and db8 < 1e-3, (N, h, rel8, db8)         # same decade as the float32 FFT ...
        assert rel8 < 5 * max(fft_rel, 5e-8), (N, h, rel8, fft_rel)
        assert rel8 < rel32                                          # ... and the reason for re-seeding every 8 frames


def test_kahan_compensation_matters_after_a_loud_passage():
    N, h, kmin, kmax, tile, T = 1024, 8, 94, 300, 512, 1200
    x = _signal(T * h, seed=h)
    want = np.fft.rfft(_frames(x, N, h, T, np.float64), axis=1)[:, kmin:kmax + 1]
    ref = np.abs(want)
    floor = 20 * np.log10(np.maximum(ref, 1e-5)) > 20 * np.log10(ref.max()) - 80

    def worst_db(y):
        mag = np.abs(y.astype(np.complex128))
        return np.abs(20 * np.log10(np.maximum(mag, 1e-5)) - 20 * np.log10(np.maximum(ref, 1e-5)))[floor].max()

    assert worst_db(_slide(x, N, h, T, kmin, kmax, tile, 8, kahan=True)) < worst_db(_slide(x, N, h, T, kmin, kmax, tile, 8, kahan=False))
