"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product (amt_tools_b200/).

CPU restatement (numpy / scipy) of the *librosa stages* that every
`amt_tools.features.*.process_audio` of the reference delegates to.  The reference
itself holds no arithmetic for this path (SURVEY.md finding 1): each module is a 1-3 line
call into librosa, which is a third-party dependency that is NOT vendored in
/root/reference and NOT installable here (requirements.txt:3 pins only `librosa>=0.9.1`;
the API usage in the reference points at librosa 0.10.x / 0.11 semantics, which is what is
restated here).  libsoxr (`res_type='soxr_hq'`, python-soxr / libsoxr 0.1.3) is likewise
absent, so the 2:1 decimator below restates soxr's *published design procedure*
(Kaiser-windowed sinc, passband 0.9136 * Nyq_out, stopband 1.0 * Nyq_out, 20-bit precision).

PARITY STATUS
  * STFT / mel filterbank / dB stages: pinned against independent implementations that are
    present in this image (torch.stft, torchaudio.functional.melscale_fbanks,
    transformers.audio_utils) -- see tests/test_oracle.py.
  * CQT / VQT / soxr stages: **parity unpinned** -- the reference ships no tests, no golden
    vectors, and librosa / soxr cannot be run here.  They are pinned only by analytic
    known-answer tests (sinusoid -> A*sqrt(L)/2) and by self-consistency.

Reference call sites restated (file:line under /root/reference/amt_tools/features):
  stft.py:66 librosa.stft            -> stft()
  mel.py:64  librosa.feature.melspectrogram -> melspectrogram(), mel_filterbank()
  vqt.py:183 librosa.vqt             -> vqt()
  common.py:199 / power.py:55 librosa.amplitude_to_db -> amplitude_to_db()
  mel.py:94  librosa.power_to_db     -> power_to_db()
  waveform.py:149 librosa.util.frame -> frame()
  common.py:254 librosa.frames_to_time -> frames_to_time()
  vqt.py:44 librosa.note_to_hz('C1') -> NOTE_C1_HZ
  vqt.py:81 librosa.cqt_frequencies  -> cqt_frequencies()
  vqt.py:87 librosa.filters.window_bandwidth('hann') -> HANN_BANDWIDTH
  vqt.py:95 librosa.core.constantq.__early_downsample_count -> early_downsample_count()
  vqt.py:219 librosa.filters.wavelet_lengths -> wavelet_lengths()
"""

import numpy as np
import scipy.fft
import scipy.signal
import scipy.sparse

# librosa.note_to_hz('C1') = 440 * 2 ** ((24 - 69) / 12)
NOTE_C1_HZ = 440.0 * 2.0 ** ((24 - 69) / 12.0)
# librosa.filters.window_bandwidth('hann') (table constant WINDOW_BANDWIDTHS['hann'])
HANN_BANDWIDTH = 1.50018310546875


# --------------------------------------------------------------------------------------
# framing / time helpers
# --------------------------------------------------------------------------------------

def frame(x, frame_length, hop_length):
    """librosa.util.frame(x, frame_length=, hop_length=) on a 1-D signal -> (frame_length, T)."""
    x = np.asarray(x)
    if x.shape[-1] < frame_length:
        raise ValueError("Input is too short (n=%d) for frame_length=%d" % (x.shape[-1], frame_length))
    n_frames = 1 + (x.shape[-1] - frame_length) // hop_length
    idx = np.arange(frame_length)[:, None] + hop_length * np.arange(n_frames)[None, :]
    return x[idx]


def frames_to_time(frames, sr, hop_length):
    """librosa.frames_to_time: samples = (asanyarray(frames) * hop).astype(int); samples / float(sr)."""
    samples = (np.asanyarray(frames) * hop_length).astype(int)
    return np.asanyarray(samples) / float(sr)


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True)."""
    if n == 1:
        return np.ones(1)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def pad_center(data, size):
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    out = np.zeros(size, dtype=data.dtype)
    out[lpad:lpad + n] = data
    return out


# --------------------------------------------------------------------------------------
# A.1  STFT
# --------------------------------------------------------------------------------------

def stft(y, n_fft=2048, hop_length=None, win_length=None, window='hann', center=True, dtype=np.float64):
    """
    librosa.stft(y, n_fft, hop_length, win_length, window, center, pad_mode='constant').
    `dtype` is the real working precision (float32 mirrors the reference, which feeds
    float32 audio and gets complex64; float64 is the "truth" used for tolerances).
    Returns complex (1 + n_fft//2, T).
    """
    y = np.asarray(y, dtype=dtype)
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 4
    if window == 'hann':
        w = hann_periodic(win_length)
    elif window == 'ones':
        w = np.ones(win_length)
    else:
        raise ValueError(window)
    w = pad_center(w, n_fft).astype(dtype)
    if center:
        y = np.pad(y, n_fft // 2, mode='constant')
    elif y.shape[-1] < n_fft:
        raise ValueError("n_fft=%d is too large for uncentered analysis of input signal of length=%d"
                         % (n_fft, y.shape[-1]))
    n_frames = 1 + (y.shape[-1] - n_fft) // hop_length
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    out = np.empty((1 + n_fft // 2, n_frames), dtype=cdtype)
    # blocked like librosa (MAX_MEM_BLOCK) so memory stays bounded on long tracks
    block = max(1, (1 << 24) // n_fft)
    for t0 in range(0, n_frames, block):
        t1 = min(n_frames, t0 + block)
        idx = np.arange(n_fft)[None, :] + hop_length * np.arange(t0, t1)[:, None]
        fr = y[idx] * w[None, :]
        out[:, t0:t1] = scipy.fft.rfft(fr, axis=-1).T
    return out


# --------------------------------------------------------------------------------------
# A.2  mel
# --------------------------------------------------------------------------------------

def hz_to_mel(f, htk=False):
    f = np.asanyarray(f, dtype=np.float64)
    if htk:
        return 2595.0 * np.log10(1.0 + f / 700.0)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def mel_to_hz(m, htk=False):
    m = np.asanyarray(m, dtype=np.float64)
    if htk:
        return 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, htk=False):
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=0, fmax=sr/2, htk=, norm='slaney', dtype=float32)."""
    if fmax is None:
        fmax = float(sr) / 2
    n_mels = int(n_mels)
    weights = np.zeros((n_mels, int(1 + n_fft // 2)), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin, htk), hz_to_mel(fmax, htk), n_mels + 2), htk)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def melspectrogram(y, sr, n_fft, hop_length, win_length, center, htk, n_mels, dtype=np.float64):
    """librosa.feature.melspectrogram(..., power=2.0): mel_basis @ |stft|**2."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length, win_length=win_length,
                    center=center, dtype=dtype)) ** 2.0
    M = mel_filterbank(sr, n_fft, n_mels=n_mels, htk=htk)
    return np.einsum('ft,mf->mt', S, M.astype(dtype), optimize=True)


# --------------------------------------------------------------------------------------
# A.4  dB
# --------------------------------------------------------------------------------------

def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S) if np.iscomplexobj(S) else S
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, magnitude))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    S = np.asarray(S)
    magnitude = np.abs(S)
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    power = np.square(magnitude, out=magnitude.copy())
    return power_to_db(power, ref=ref_value ** 2, amin=amin ** 2, top_db=top_db)


# --------------------------------------------------------------------------------------
# soxr 'HQ' 2^k:1 decimator -- restated design procedure (libsoxr 0.1.3: soxr.c
# soxr_quality_spec, cr.c _soxr_init, filter.c lsx_design_lpf / lsx_kaiser_beta /
# lsx_kaiser_params / lsx_make_lpf).  Written from memory of the published source; the
# constants below are the ones that procedure uses.  PARITY UNPINNED (no libsoxr here).
# --------------------------------------------------------------------------------------

_KAISER_COEFS = np.array([
    [-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + .001],
    [-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + .002],
    [-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + .003],
    [-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + .006],
    [8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + .015],
    [9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + .025],
    [-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + .05],
    [-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + .085],
    [1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + .1],
    [-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + .18],
])


def _kaiser_beta(att, tr_bw):
    if att >= 60:
        realm = np.log(tr_bw / .0005) / np.log(2.)
        i0 = int(np.clip(int(realm), 0, len(_KAISER_COEFS) - 1))
        i1 = int(np.clip(1 + int(realm), 0, len(_KAISER_COEFS) - 1))
        c0, c1 = _KAISER_COEFS[i0], _KAISER_COEFS[i1]
        b0 = ((c0[0] * att + c0[1]) * att + c0[2]) * att + c0[3]
        b1 = ((c1[0] * att + c1[1]) * att + c1[2]) * att + c1[3]
        return b0 + (b1 - b0) * (realm - int(realm))
    if att > 50:
        return .1102 * (att - 8.7)
    if att > 20.96:
        return .58417 * (att - 20.96) ** .4 + .07886 * (att - 20.96)
    return 0.0


def soxr_hq_taps(factor=2):
    """
    Linear-phase low-pass for an integer `factor`:1 decimation at soxr 'HQ' quality.
    HQ: precision 20 bit -> rej = 20*20*log10(2); passband_end = 1 - .05/TO_3dB(rej),
    stopband_begin = 1 (both relative to the OUTPUT Nyquist); att = (20+1)*20*log10(2).
    DC gain 1.
    """
    bits = 20.0
    rej = bits * 20.0 * np.log10(2.0)
    to3db = (1.6e-6 * rej - 7.5e-4) * rej + .646
    Fp0 = 1.0 - .05 / to3db
    Fs0 = 1.0
    att = (bits + 1) * 20.0 * np.log10(2.0)
    Fn = float(factor)
    Fp, Fs = Fp0 / Fn, Fs0 / Fn
    tr_bw = .5 * (Fs - Fp)
    tr_bw = min(tr_bw, .5 * Fs)
    Fc = Fs - tr_bw
    beta = _kaiser_beta(att, tr_bw * .5 / Fc)
    a = ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902
    num_taps = int(np.ceil(a / tr_bw + 1))
    modulo = 4  # k = -4: num_taps = 1 (mod 4)
    num_taps = (num_taps + modulo - 2) // modulo * modulo + 1
    m = num_taps - 1
    rho = .5
    mult1 = 1.0 / (.5 * m + rho)
    z = np.arange(num_taps) - .5 * m
    x = z * np.pi
    with np.errstate(invalid='ignore', divide='ignore'):
        h = np.where(x != 0, np.sin(Fc * x) / x, Fc)
    yv = z * mult1
    h = h * np.i0(beta * np.sqrt(1 - yv * yv)) / np.i0(beta)
    return h / h.sum()


_TAPS_CACHE = {}


def resample_decimate(y, factor, dtype=np.float64):
    """
    librosa.resample(y, orig_sr=factor, target_sr=1, res_type='soxr_hq', scale=True):
    zero-phase (delay-compensated) FIR decimation, output length ceil(n/factor),
    then y_hat /= sqrt(ratio) with ratio = 1/factor (i.e. multiply by sqrt(factor)).
    """
    if factor not in _TAPS_CACHE:
        _TAPS_CACHE[factor] = soxr_hq_taps(factor)
    h = _TAPS_CACHE[factor].astype(dtype)
    y = np.asarray(y, dtype=dtype)
    n_out = int(np.ceil(y.shape[-1] / float(factor)))
    D = (len(h) - 1) // 2
    full = scipy.signal.upfirdn(h, np.concatenate([y, np.zeros(D + factor, dtype=dtype)]), up=1, down=1)
    out = full[D:D + factor * n_out:factor][:n_out]
    return (out * np.sqrt(float(factor))).astype(dtype)


# --------------------------------------------------------------------------------------
# A.3  VQT / CQT
# --------------------------------------------------------------------------------------

def cqt_frequencies(n_bins, fmin, bins_per_octave=12):
    return fmin * 2.0 ** (np.arange(0, n_bins, dtype=float) / bins_per_octave)


def relative_bandwidth_et(bins_per_octave):
    """librosa >= 0.10 alpha for equal-tempered spacing: (r^2 - 1) / (r^2 + 1), r = 2^(1/bpo)."""
    r = 2.0 ** (1.0 / bins_per_octave)
    return (r ** 2 - 1) / (r ** 2 + 1)


def wavelet_lengths(freqs, sr, gamma, alpha, filter_scale=1.0):
    freqs = np.atleast_1d(np.asarray(freqs, dtype=float))
    Q = float(filter_scale) / alpha
    f_cutoff = np.max(freqs * (1 + 0.5 * HANN_BANDWIDTH / Q) + 0.5 * gamma)
    lengths = Q * sr / (freqs + gamma / alpha)
    return lengths, f_cutoff


def wavelet(freqs, sr, gamma, alpha):
    """librosa.filters.wavelet(freqs=, sr=, window='hann', filter_scale=1, pad_fft=True, norm=1, gamma=, alpha=)."""
    lengths, _ = wavelet_lengths(freqs, sr, gamma, alpha)
    filters = []
    for ilen, freq in zip(lengths, freqs):
        n = np.arange(-ilen // 2, ilen // 2, dtype=float)
        sig = np.exp(1j * (n * 2 * np.pi * freq / sr))
        sig = sig * hann_periodic(len(sig))
        sig = sig / np.sum(np.abs(sig))
        filters.append(sig)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    filters = np.asarray([pad_center(f, max_len) for f in filters], dtype=np.complex64)
    return filters, lengths


def sparsify_rows(x, quantile=0.01):
    """librosa.util.sparsify_rows -> CSR (same dtype as x)."""
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    x_sparse = scipy.sparse.lil_matrix(x.shape, dtype=x.dtype)
    for i, j in enumerate(threshold_idx):
        idx = np.where(mags[i] >= mag_sort[i, j])
        x_sparse[i, idx] = x[i, idx]
    return x_sparse.tocsr()


def vqt_filter_fft(sr, freqs, gamma, alpha, sparsity=0.01):
    """librosa.core.constantq.__vqt_filter_fft (hop_length=None): complex64 CSR basis, n_fft, lengths."""
    basis, lengths = wavelet(freqs, sr, gamma, alpha)
    n_fft = basis.shape[1]
    basis *= (lengths[:, np.newaxis] / float(n_fft))
    fft_basis = scipy.fft.fft(basis, n=n_fft, axis=1)[:, :(n_fft // 2) + 1]
    fft_basis = sparsify_rows(fft_basis, quantile=sparsity)
    return fft_basis, n_fft, lengths


def num_two_factors(x):
    if x <= 0:
        return 0
    n = 0
    while x % 2 == 0:
        n += 1
        x //= 2
    return n


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves):
    c1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    c2 = max(0, num_two_factors(hop_length) - n_octaves + 1)
    return min(c1, c2)


def vqt(y, sr=22050, hop_length=512, fmin=None, n_bins=84, bins_per_octave=12, gamma=None,
        sparsity=0.01, dtype=np.float64, basis_cache=None):
    """
    librosa.vqt(y, sr=, hop_length=, fmin=, n_bins=, bins_per_octave=, gamma=) with all other
    arguments at their defaults (filter_scale=1, norm=1, sparsity=0.01, window='hann',
    scale=True, pad_mode='constant', res_type='soxr_hq', tuning=0.0).  Returns complex (n_bins, T).
    `basis_cache` (dict) optionally memoises the per-octave bases (librosa rebuilds them on
    every call; the cached variant is the second CPU-baseline flavour of SURVEY.md 8d).
    """
    y = np.asarray(y, dtype=dtype)
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    if fmin is None:
        fmin = NOTE_C1_HZ
    freqs = cqt_frequencies(n_bins, fmin, bins_per_octave)
    alpha = relative_bandwidth_et(bins_per_octave)
    if gamma is None:
        gamma = alpha * 24.7 / 0.108
    lengths, filter_cutoff = wavelet_lengths(freqs, sr, gamma, alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ValueError("Wavelet basis with max frequency=%g would exceed the Nyquist frequency=%g"
                         % (np.max(freqs), nyquist))
    if num_two_factors(hop_length) < n_octaves - 1:
        raise ValueError("hop_length must be a positive integer multiple of 2^%d for %d-octave CQT/VQT"
                         % (n_octaves - 1, n_octaves))
    # early downsampling
    eds = early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves)
    if eds > 0:
        factor = 2 ** eds
        hop_length //= factor
        if y.shape[-1] < factor:
            raise ValueError("Input signal length=%d is too short for %d-octave CQT" % (len(y), n_octaves))
        y = resample_decimate(y, factor, dtype=dtype)
        sr = sr / float(factor)
    cdtype = np.complex64 if dtype == np.float32 else np.complex128
    my_y, my_sr, my_hop = y, sr, hop_length
    resp = []
    for i in range(n_octaves):
        lo = max(0, n_bins - n_filters * (i + 1))
        hi = n_bins - n_filters * i
        key = (float(my_sr), float(freqs[lo]), hi - lo, float(gamma), int(bins_per_octave))
        if basis_cache is not None and key in basis_cache:
            fft_basis, n_fft = basis_cache[key]
        else:
            fft_basis, n_fft, _ = vqt_filter_fft(my_sr, freqs[lo:hi], gamma, alpha, sparsity)
            if basis_cache is not None:
                basis_cache[key] = (fft_basis, n_fft)
        fb = fft_basis * np.sqrt(sr / my_sr)
        D = stft(my_y, n_fft=n_fft, hop_length=my_hop, window='ones', center=True, dtype=dtype)
        resp.append(fb.astype(cdtype).dot(D))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            my_y = resample_decimate(my_y, 2, dtype=dtype)
    # __trim_stack
    max_col = min(c.shape[-1] for c in resp)
    V = np.empty((n_bins, max_col), dtype=cdtype)
    end = n_bins
    for c in resp:
        n_oct = c.shape[0]
        if end < n_oct:
            V[:end, :] = c[-end:, :max_col]
        else:
            V[end - n_oct:end, :] = c[:, :max_col]
        end -= n_oct
    lengths, _ = wavelet_lengths(freqs, sr, gamma, alpha)
    V /= np.sqrt(lengths)[:, None].astype(dtype)
    return V
