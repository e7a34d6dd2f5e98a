"""
CPU test of the exact-ladder geometry (csrc/host_plan.cpp clip_tail_layout) against the oracle.

A harmonic with eds >= 2 is early-downsampled by librosa in ONE resample call (features/vqt.py:183 -> librosa.vqt ->
__early_downsample); the CUDA path serves its interior frames from the shared cascaded 2:1 ladder and recomputes only
the tail frames from an "exact ladder" whose levels are built from sample `first` on.  Here the same scheme is
restated in float64 numpy with the numbers the C-ABI reports (amtfeat_clip_describe), every sample the scheme must not
read is poisoned with NaN, and the stitched result is compared with the oracle's one-shot path.
"""
import numpy as np
import pytest

import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
from oracle import librosa_stages as ls


def fir_tail(src, first_in, len_in, taps, factor, first_out, len_out):
    """y[m] = sum_k taps[k] * src[factor * m + D - k] for first_out <= m < len_out, zero extension outside [0, len_in);
    `src` is NaN below first_in: any read of a sample the scheme does not hold poisons the result."""
    D = (len(taps) - 1) // 2
    out = np.full(len_out, np.nan)
    k = np.arange(len(taps))
    for m0 in range(first_out, len_out, 2048):
        m = np.arange(m0, min(len_out, m0 + 2048))
        idx = factor * m[:, None] + D - k[None, :]
        ok = (idx >= 0) & (idx < len_in)
        vals = np.where(ok, src[np.clip(idx, 0, len_in - 1)], 0.0)
        out[m] = vals @ taps
    return out


def octave_response(sig, level, i, cfg, sr_post):
    """Response of octave i (from the top) of the harmonic on the level signal `sig` (librosa.vqt's loop body)."""
    sr, hop, fmin, n_bins, bpo = cfg
    freqs = ls.cqt_frequencies(n_bins, fmin, bpo)
    alpha = ls.relative_bandwidth_et(bpo)
    nf = min(bpo, n_bins)
    lo, hi = max(0, n_bins - nf * (i + 1)), n_bins - nf * i
    my_sr = sr / 2.0 ** level
    fb, n_fft, _ = ls.vqt_filter_fft(my_sr, freqs[lo:hi], 0.0, alpha)
    fb = fb * np.sqrt(sr_post / my_sr)
    D = ls.stft(sig, n_fft=n_fft, hop_length=hop >> level, window='ones', center=True, dtype=np.float64)
    return fb.astype(np.complex128).dot(D), n_fft


@pytest.mark.parametrize('n,n_bins,eds', [(22050 * 5 + 137, 144, 2), (70001, 144, 2), (22050 * 4 + 1, 120, 3)])
def test_tail_scheme_equals_one_shot_early_downsampling(n, n_bins, eds):
    sr, hop, bpo, h = 22050, 256, 24, 0.5
    fmin = h * ls.NOTE_C1_HZ
    factor = 2 ** eds
    m = ab.HCQT(sr, hop, False, harmonics=[h], n_bins=n_bins, bins_per_octave=bpo)
    plan = m.describe()
    assert plan['eds_lib'] == [eds] and plan['exact_ladders'] == [{'eds': eds, 'taps': len(ls.soxr_hq_taps(factor))}]
    lay = m.describe_clip(n)
    n_oct = plan['n_octaves']
    D = lay['decim_delay']
    h2 = ls.soxr_hq_taps(2) * np.sqrt(2.0)
    h4 = ls.soxr_hq_taps(factor) * np.sqrt(float(factor))
    assert D == (len(h2) - 1) // 2
    # piano-like audio plus a steady low tone, so that the clip starts and ends abruptly inside the harmonic's band
    x = piano_like(n, sr, seed=5).astype(np.float64) + 0.5 * np.sin(2 * np.pi * 61.0 * np.arange(n) / sr + 1.0)
    lens = lay['level_len']
    # shared cascade (what librosa does between octaves) and the exact chain, both complete
    shared = [x]
    for l in range(1, eds + n_oct):
        shared.append(ls.resample_decimate(shared[-1], 2))
        assert len(shared[-1]) == lens[l]
    exact = {eds: ls.resample_decimate(x, factor)}
    for l in range(eds + 1, eds + n_oct):
        exact[l] = ls.resample_decimate(exact[l - 1], 2)
    # the head and tail pieces of the exact chain, built the way the kernels build them, from poisoned storage
    first, count, head = (lay['exact_ladders'][0][k] for k in ('first', 'count', 'head'))
    piece = {}
    for l in range(eds, eds + n_oct):
        assert first[l] >= 0 and first[l] % 4 == 0 and count[l] == lens[l] - first[l]
        hl = head[l] if head[l] >= 0 else 0
        assert (head[l] == -1 and first[l] == 0) or (0 <= head[l] <= first[l] and head[l] % 4 == 0) or head[l] == lens[l]
        src, src_len, taps, f = (x, n, h4, factor) if l == eds else (piece[l - 1], lens[l - 1], h2, 2)
        piece[l] = fir_tail(src, 0, src_len, taps, f, first[l], lens[l])
        if hl:
            piece[l][:hl] = fir_tail(src, 0, src_len, taps, f, 0, hl)[:hl]
        stored = np.r_[0:hl, first[l]:lens[l]]
        assert np.all(np.isfinite(piece[l][stored])), 'level %d reads samples of level %d that are not stored' % (l, l - 1)
        np.testing.assert_allclose(piece[l][stored], exact[l][stored], rtol=0, atol=1e-12 * np.abs(exact[l]).max())
    cfg = (sr, hop, fmin, n_bins, bpo)
    sr_post = sr / float(factor)
    resp, worst_interior = [], 0.0
    for i in range(n_oct):
        l = eds + i
        t0, th = lay['alt_t0'][l], lay['alt_th'][l]
        a, n_fft = octave_response(shared[l], l, i, cfg, sr_post)
        b, _ = octave_response(piece[l], l, i, cfg, sr_post)
        c, _ = octave_response(exact[l], l, i, cfg, sr_post)
        T_l = a.shape[1]
        assert lay['alt_nfft_max'][l] == n_fft
        if t0 < 0:
            t0 = T_l            # no frame of this level reaches the part of the shared signal that deviates
        A = max(32, 16384 // n_fft)
        assert t0 % A == 0 and th % A == 0 and th <= t0
        # frames in [th, t0) read only shared samples inside [hsafe, dev); the others only stored samples of the exact pieces
        if th < t0:
            assert th * (hop >> l) - n_fft // 2 >= lay['hsafe'][l] and (t0 - 1) * (hop >> l) + n_fft // 2 <= lay['dev'][l]
        assert np.all(np.isfinite(b[:, :th])) and np.all(np.isfinite(b[:, t0:]))
        scale = np.abs(c).max()
        if th < t0:
            worst_interior = max(worst_interior, np.abs(a[:, th:t0] - c[:, th:t0]).max() / scale)
        assert np.abs(b[:, :th] - c[:, :th]).max(initial=0) <= 1e-12 * scale and np.abs(b[:, t0:] - c[:, t0:]).max(initial=0) <= 1e-12 * scale
        resp.append(np.concatenate([b[:, :th], a[:, th:t0], b[:, t0:]], axis=1))
    # pass band of the cascade == pass band of the one-shot filter: the interior agrees to filter ripple
    assert worst_interior < 5e-7, worst_interior
    T = min(r.shape[1] for r in resp)
    V = np.concatenate([r[:, :T] for r in resp[::-1]], axis=0)
    lengths, _ = ls.wavelet_lengths(ls.cqt_frequencies(n_bins, fmin, bpo), sr_post, 0.0, ls.relative_bandwidth_et(bpo))
    V = np.abs(V / np.sqrt(lengths)[:, None])
    want = np.abs(ls.vqt(x, sr=sr, hop_length=hop, fmin=fmin, n_bins=n_bins, bins_per_octave=bpo, gamma=0.0))
    assert V.shape == want.shape
    assert np.linalg.norm(V - want) / np.linalg.norm(want) < 1e-6
    # and the deviation this removes is real: the shared ladder alone is off in the last frames
    plain = np.concatenate([octave_response(shared[eds + i], eds + i, i, cfg, sr_post)[0][:, :T] for i in range(n_oct)][::-1], axis=0)
    plain = np.abs(plain / np.sqrt(lengths)[:, None])
    peak = want.max()
    assert np.abs(V - want).max() < 1e-6 * peak
    # (a few 1e-6 of the peak: it shows in bins 80 dB down, which is where the dB features of a quiet passage live)
    assert np.abs(plain - want)[:, -8:].max() > 10 * np.abs(V - want).max()
    assert np.abs(plain - want)[:, :8].max() > 3 * np.abs(V - want).max()


def test_exact_ladders_only_where_librosa_downsamples_in_one_shot():
    std = ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60)
    d = std.describe()
    assert d['eds_lib'] == [2, 1, 0, 0, 0, 0] and d['alt_mask'] == 1 and len(d['exact_ladders']) == 1
    assert sum(it['alt'] for it in d['items']) == 6          # h = 0.5: six octaves on levels 2 .. 7
    assert ab.CQT(22050, 512, n_bins=192, bins_per_octave=24).describe()['exact_ladders'] == []
    assert ab.HVQT(22050, 512, harmonics=[1, 2, 3], n_bins=72, bins_per_octave=12).describe()['exact_ladders'] == []
    two = ab.HCQT(22050, 512, harmonics=[0.25, 0.5, 1], n_bins=96, bins_per_octave=24).describe()
    assert [a['eds'] for a in two['exact_ladders']] == sorted(set(e for e in two['eds_lib'] if e >= 2), reverse=True) or \
        sorted(a['eds'] for a in two['exact_ladders']) == sorted(set(e for e in two['eds_lib'] if e >= 2))
    # a clip shorter than the decimator: everything is "tail"
    lay = std.describe_clip(3000)
    assert all(t == 0 for t in lay['alt_t0'][2:]) and all(t == 0 for t in lay['alt_th'][2:]) and lay['exact_ladders'][0]['first'][2:] == [0] * 6


def test_db_maximum_runs_over_each_harmonics_own_frames():
    # hvqt.py:123-128: every harmonic is converted to dB over its own VQT, then trimmed to the common frame count
    m = ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60)
    lin = ab.HCQT(22050, 256, False, n_bins=360, bins_per_octave=60)
    seen_longer = 0
    for n in list(range(22050, 22050 + 1024, 37)) + [22526, 23037, 23807]:
        lay = m.describe_clip(n)
        assert lay['frames'] == m.get_expected_frames(np.empty(n, np.float32)) == min(lay['harmonic_frames'])
        assert lay['frames_computed'] == max(lay['harmonic_frames'])
        seen_longer += lay['frames_computed'] > lay['frames']
        assert lin.describe_clip(n)['frames_computed'] == lay['frames']      # no dB: nothing past the stored frames is needed
    assert seen_longer >= 3
    assert m.describe_clip(23807)['harmonic_frames'] == [94, 94, 93, 93, 93, 93]
