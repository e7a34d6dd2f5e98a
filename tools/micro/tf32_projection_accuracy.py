"""
CPU emulation: what a 3xTF32 tensor-core projection would do to the dB features (no GPU needed).

The wavelet-basis projection out[r, t] = sum_k W[r, k] D[k, t] (librosa.vqt's `fft_basis.dot(D)`, features/vqt.py:183) is
computed for one real HCQT item (level, n_fft) of a synthetic clip in
  * float64                                  (truth)
  * float32 operands, float32 accumulation   (what the FFMA2 projection of cqt_kernel does)
  * 3xTF32: W = Wh + Wl, D = Dh + Dl with TF32 pieces (10 explicit mantissa bits), Wh Dh + Wh Dl + Wl Dh, float32 accumulation
    (the split tcgen05.mma.kind::tf32 would need; Wl Dl dropped, pieces rounded to nearest -- the favourable variant)
  * 1xTF32                                   (for scale)
and the resulting dB features are compared on the bins within 60 dB of the maximum (the north_star bar: 1e-3 dB).

    python tools/micro/tf32_projection_accuracy.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import librosa_stages as ls  # noqa: E402


def tf32(x):
    """Round float32 to TF32 (1 + 8 + 10 bits), round to nearest even."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def split(x):
    hi = tf32(x)
    lo = tf32(np.asarray(x, np.float32) - hi)
    return hi, lo


def cmatmul32(Wr, Wi, Dr, Di):
    """Complex product from four real float32 GEMMs (float32 accumulation)."""
    f = np.float32
    return (Wr.astype(f) @ Dr.astype(f) - Wi.astype(f) @ Di.astype(f)), (Wr.astype(f) @ Di.astype(f) + Wi.astype(f) @ Dr.astype(f))


def main():
    sr, hop, bpo, n_bins = 22050, 256, 60, 360
    y = piano_like(sr * 10, sr, seed=3).astype(np.float64)
    freqs = ls.cqt_frequencies(n_bins, ls.NOTE_C1_HZ, bpo)
    alpha = ls.relative_bandwidth_et(bpo)
    out = {}
    for level in (0, 3):           # top octave of h = 1 at full rate, and a decimated one
        sig = y
        for _ in range(level):
            sig = ls.resample_decimate(sig, 2)
        lo_, hi_ = n_bins - bpo * (level + 1), n_bins - bpo * level
        fb, n_fft, _ = ls.vqt_filter_fft(sr / 2.0 ** level, freqs[lo_:hi_], 0.0, alpha)
        W = fb.toarray().astype(np.complex64)                      # librosa keeps the basis in complex64
        D = ls.stft(sig.astype(np.float32), n_fft=n_fft, hop_length=hop >> level, window='ones', center=True, dtype=np.float32)
        truth = W.astype(np.complex128) @ D.astype(np.complex128)
        Wr, Wi, Dr, Di = W.real, W.imag, D.real, D.imag
        r32, i32 = cmatmul32(Wr, Wi, Dr, Di)
        (Wrh, Wrl), (Wih, Wil), (Drh, Drl), (Dih, Dil) = split(Wr), split(Wi), split(Dr), split(Di)
        r3 = i3 = 0
        for (a_r, a_i), (b_r, b_i) in (((Wrh, Wih), (Drh, Dih)), ((Wrh, Wih), (Drl, Dil)), ((Wrl, Wil), (Drh, Dih))):
            rr, ii = cmatmul32(a_r, a_i, b_r, b_i)
            r3, i3 = r3 + rr, i3 + ii
        r1, i1 = cmatmul32(Wrh, Wih, Drh, Dih)

        def db(re, im):
            p = np.asarray(re, np.float64) ** 2 + np.asarray(im, np.float64) ** 2
            return 10 * np.log10(np.maximum(p, 1e-10))
        ref = db(truth.real, truth.imag)
        ref -= ref.max()
        top = ref > -60
        res = {}
        for name, (re, im) in (('fp32', (r32, i32)), ('3xTF32', (r3, i3)), ('1xTF32', (r1, i1))):
            d = db(re, im)
            d -= d.max()
            res[name] = {'dB_maxabs_above_-60dB': float(np.abs(d - ref)[top].max()),
                         'rel_l2': float(np.linalg.norm((re + 1j * im) - truth) / np.linalg.norm(truth))}
        out['level%d_nfft%d' % (level, n_fft)] = res
        print('level', level, 'n_fft', n_fft, json.dumps(res))
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02_tf32_projection_accuracy.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
