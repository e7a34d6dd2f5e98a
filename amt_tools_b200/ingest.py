"""
Audio ingest on the device -- the step before the feature modules.

Mirrors `tools.load_normalize_audio` (/root/reference/amt_tools/tools/io.py:50-87) for everything that follows the file
decoder: `librosa.load(..., sr=fs, mono=True, res_type=...)` = channel mean + `librosa.resample` (resampy's windowed-sinc
interpolation for res_type 'kaiser_best' / 'kaiser_fast'), then `tools.rms_norm` (amt_tools/tools/utils.py:2789-2814).
Decoded samples go in (NumPy or torch, host or device), normalised mono audio at the target rate comes out as a CUDA
tensor that the feature modules consume without a host round trip.  All arithmetic runs in libamtfeat.so
(csrc/ingest.cu); there is no CPU fallback.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib

_FILTERS = {'kaiser_best': _lib.RES_KAISER_BEST, 'kaiser_fast': _lib.RES_KAISER_FAST}


def _device(device):
    dev = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0) if device is None \
        else torch.device(device)
    if dev.type != 'cuda':
        raise ValueError('amt_tools_b200 computes on CUDA devices only (no CPU fallback)')
    if dev.index is None:
        dev = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
    return dev


def _to_device(audio, dev):
    t = torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)) if isinstance(audio, np.ndarray) else audio
    return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _workspace(batch, dev):
    nbytes = int(_lib.lib.amtfeat_ingest_workspace_bytes(batch))
    return torch.empty(nbytes, dtype=torch.uint8, device=dev), nbytes


class Resampler(object):
    """One (orig_sr, target_sr, res_type) interpolation table.  `device=None` with no GPU gives a host-only object
    (output lengths and the table, no compute)."""

    def __init__(self, orig_sr, target_sr, res_type='kaiser_best', device=None, host_only=False):
        if res_type not in _FILTERS:
            raise ValueError("res_type must be 'kaiser_best' or 'kaiser_fast' (resampy filters; soxr / polyphase modes "
                             "of librosa.resample are not implemented)")
        self.orig_sr, self.target_sr, self.res_type = orig_sr, target_sr, res_type
        self.device = None if host_only else _device(device)
        self.handle = C.c_void_p()
        _lib.check(_lib.lib.amtfeat_resampler_create(float(orig_sr), float(target_sr), _FILTERS[res_type],
                                                     -1 if host_only else self.device.index, C.byref(self.handle)))

    def __del__(self):
        if getattr(self, 'handle', None) and _lib is not None and getattr(_lib, 'lib', None) is not None:   # (module teardown)
            _lib.lib.amtfeat_resampler_destroy(self.handle)
            self.handle = None

    def out_len(self, n):
        return int(_lib.lib.amtfeat_resampler_out_len(self.handle, int(n)))

    def table(self):
        nt, step = C.c_int(), C.c_int()
        n = int(_lib.lib.amtfeat_resampler_table(self.handle, None, 0, C.byref(nt), C.byref(step)))
        win = np.empty(n, dtype=np.float64)
        _lib.lib.amtfeat_resampler_table(self.handle, win.ctypes.data_as(C.POINTER(C.c_double)), n, None, None)
        return win, nt.value, step.value

    def __call__(self, audio):
        """Resample one mono clip (N,) or a list of clips; returns CUDA tensors of length ceil(N * ratio) (librosa.resample, fix=True)."""
        if self.device is None:
            raise _lib.AmtfeatError('host-only resampler: no CUDA device (there is no CPU compute path)')
        single = not isinstance(audio, (list, tuple))
        clips = [_to_device(a, self.device) for a in ([audio] if single else audio)]
        for c in clips:
            if c.ndim != 1:
                raise ValueError('resample expects mono clips of shape (N,)')
        n_in = [int(c.numel()) for c in clips]
        n_out = [self.out_len(n) for n in n_in]
        in_off, out_off, a, b = [], [], 0, 0
        for ni, no in zip(n_in, n_out):
            in_off.append(a)
            out_off.append(b)
            a += (ni + 3) // 4 * 4
            b += (no + 3) // 4 * 4
        buf = clips[0] if len(clips) == 1 else torch.zeros(max(a, 1), dtype=torch.float32, device=self.device)
        if len(clips) > 1:
            for c, o in zip(clips, in_off):
                buf[o:o + c.numel()] = c
        out = torch.empty(max(b, 1), dtype=torch.float32, device=self.device)
        ws, ws_bytes = _workspace(len(clips), self.device)
        _lib.check(_lib.lib.amtfeat_resample(self.handle, buf.data_ptr(), _lib.i64_array(in_off), _lib.i64_array(n_in),
                                             len(clips), out.data_ptr(), _lib.i64_array(out_off), ws.data_ptr(), ws_bytes,
                                             _stream(self.device)))
        outs = [out[o:o + n] for o, n in zip(out_off, n_out)]
        return outs[0] if single else outs


_RESAMPLERS = {}


def resample(audio, orig_sr, target_sr, res_type='kaiser_best', device=None):
    """librosa.resample(y, orig_sr=, target_sr=, res_type=) for the resampy filters, on the device."""
    dev = _device(device)
    key = (float(orig_sr), float(target_sr), res_type, dev.index)
    if key not in _RESAMPLERS:
        _RESAMPLERS[key] = Resampler(orig_sr, target_sr, res_type, dev)
    return _RESAMPLERS[key](audio)


def pcm16_to_float(pcm, scale=1.0 / 32768.0, device=None):
    """
    16-bit PCM samples (np.int16 array or torch.int16 tensor, any shape, host or device) -> float32 CUDA tensor, pcm * scale:
    what soundfile / librosa.load (tools/io.py:78) do on the host, after an upload of half the bytes.
    """
    dev = _device(device)
    t = pcm if isinstance(pcm, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pcm, dtype=np.int16))
    if t.dtype != torch.int16:
        raise ValueError('pcm16_to_float expects int16 samples')
    t = t.to(dev, non_blocking=True).contiguous()
    out = torch.empty(t.shape, dtype=torch.float32, device=dev)
    if t.numel():
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.amtfeat_pcm16_to_float(t.data_ptr(), t.numel(), float(scale), out.data_ptr(), _stream(dev)))
    return out


def to_mono(audio, device=None):
    """librosa.to_mono: (channels, N) -> (N,), the mean over channels; mono input is passed through."""
    dev = _device(device)
    x = _to_device(audio, dev)
    if x.ndim == 1:
        return x
    if x.ndim != 2:
        raise ValueError('to_mono expects (channels, N) or (N,)')
    ch, n = int(x.shape[0]), int(x.shape[1])
    out = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib.amtfeat_to_mono(x.data_ptr(), n, ch, out.data_ptr(), _stream(dev)))
    return out


def rms_norm(audio, device=None):
    """tools.rms_norm (utils.py:2789-2814): audio / sqrt(mean(audio ** 2)); silent audio is returned unchanged."""
    dev = _device(device)
    x = _to_device(audio, dev)
    if x.ndim != 1:
        raise ValueError('rms_norm expects mono audio of shape (N,)')
    if isinstance(audio, torch.Tensor) and x.data_ptr() == audio.data_ptr():
        x = x.clone()   # the reference returns a new array
    ws, ws_bytes = _workspace(1, dev)
    _lib.check(_lib.lib.amtfeat_rms_norm(x.data_ptr(), _lib.i64_array([0]), _lib.i64_array([x.numel()]), 1, ws.data_ptr(),
                                         ws_bytes, _stream(dev)))
    return x


def load_normalize_audio(audio, orig_sr, fs=None, norm=-1, res_type='kaiser_best', device=None):
    """
    io.py:50-87 after the decoder: `audio` holds the decoded samples ((channels, N) or (N,)) at `orig_sr`.
    Returns (audio at `fs`, mono, RMS-normalised when norm == -1; fs).  Only the reference's default normalisation (-1) and
    None are provided; librosa.util.normalize norms are not part of the hot path.
    """
    if norm not in (-1, None):
        raise ValueError('norm must be -1 (root-mean-square) or None')
    x = to_mono(audio, device)
    if fs is None:
        fs = orig_sr
    if float(fs) != float(orig_sr):
        x = resample(x, orig_sr, fs, res_type, device)
    if norm == -1:
        x = rms_norm(x, device)
    return x, fs
