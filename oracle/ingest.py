"""
TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

CPU restatement of the audio ingest path of the reference:
  tools.load_normalize_audio   /root/reference/amt_tools/tools/io.py:50-87
  tools.rms_norm               /root/reference/amt_tools/tools/utils.py:2789-2814
The arithmetic of `librosa.load(sr=fs, mono=True, res_type='kaiser_best')` lives in third-party code that is NOT in this
image: librosa (unpinned, `librosa>=0.9.1`, requirements.txt:3) -> resampy (0.4.x) for the 'kaiser_*' filters.  resampy's
published algorithm is restated here: `filters.sinc_window` (Kaiser-windowed sinc, 2**precision table entries per zero
crossing) and the interpolation loop of `interpn._resample_loop` (left / right wing, linear interpolation between table
entries, table stride int(scale * num_table)).  PARITY UNPINNED against resampy itself: no resampy run, golden vector or test of
the reference is available here.  What pins the restatement (tests/test_ingest.py): known answers (DC gain, in-band sinusoid,
output length) and torchaudio's Kaiser-windowed sinc resampler with its documented 'kaiser_best' equivalents, an independent
implementation that agrees to 1e-6 of the peak on in-band content wherever resampy's table stride is exact (2 : 1, upsampling);
the transition band and the stride truncation of the other ratios stay recall-only.  rms_norm / to_mono follow the reference's own lines.
"""

import numpy as np
import scipy.signal

FILTERS = {
    # name: (num_zeros, precision, kaiser beta, rolloff) -- resampy/filters.py documentation of the shipped tables
    'kaiser_best': (64, 9, 14.769656459379492, 0.9475937167399596),
    'kaiser_fast': (16, 9, 8.555504641634386, 0.85),
}


def sinc_window(num_zeros, precision, beta, rolloff):
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = scipy.signal.get_window(('kaiser', beta), 2 * n + 1, fftbins=False)[n:]
    return taper * sinc_win, num_bits, rolloff


def resample(x, sr_orig, sr_new, res_type='kaiser_best'):
    x = np.asarray(x)
    ratio = float(sr_new) / sr_orig
    n_out = int(x.shape[-1] * ratio)                # resampy.resample: shape[axis] = int(shape[axis] * sample_ratio)
    n_fix = int(np.ceil(x.shape[-1] * ratio))       # librosa.resample(fix=True): util.fix_length(y_hat, size=ceil(n * ratio))
    interp_win, num_table, _ = sinc_window(*FILTERS[res_type])
    if ratio < 1:
        interp_win = ratio * interp_win
    interp_delta = np.diff(interp_win, append=interp_win[-1])
    scale = min(1.0, ratio)
    index_step = int(scale * num_table)
    nwin, n_orig = interp_win.shape[0], x.shape[-1]
    y = np.zeros(n_out, dtype=np.float64)
    xd = x.astype(np.float64)
    t_out = np.arange(n_out) * (1.0 / ratio)
    for t in range(n_out):
        time_register = t_out[t]
        n = int(time_register)
        frac = scale * (time_register - n)
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        i_max = min(n + 1, (nwin - offset) // index_step)
        idx = offset + np.arange(i_max) * index_step
        y[t] += np.dot(interp_win[idx] + eta * interp_delta[idx], xd[n - np.arange(i_max)])
        frac = scale - frac
        index_frac = frac * num_table
        offset = int(index_frac)
        eta = index_frac - offset
        k_max = min(n_orig - n - 1, (nwin - offset) // index_step)
        idx = offset + np.arange(k_max) * index_step
        y[t] += np.dot(interp_win[idx] + eta * interp_delta[idx], xd[n + 1 + np.arange(k_max)])
    y = np.concatenate([y, np.zeros(n_fix - n_out)])     # fix_length pads with zeros
    return y.astype(x.dtype if x.dtype.kind == 'f' else np.float64)


def to_mono(y):
    y = np.asarray(y)
    return y if y.ndim == 1 else np.mean(y, axis=tuple(range(y.ndim - 1)))   # librosa.to_mono


def rms_norm(audio):
    rms = np.sqrt(np.mean(audio ** 2))      # utils.py:2807
    if rms > 0:                              # utils.py:2810
        audio = audio / rms
    return audio


def load_normalize_audio(samples, orig_sr, fs=None, norm=-1, res_type='kaiser_best'):
    audio = to_mono(np.asarray(samples, dtype=np.float32))
    if fs is not None and fs != orig_sr:
        audio = resample(audio, orig_sr, fs, res_type).astype(np.float32)
    else:
        fs = orig_sr
    if norm == -1:
        audio = rms_norm(audio)
    return audio, fs
