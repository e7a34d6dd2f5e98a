"""
Host-side mirror of `amt_tools.features` over libamtfeat.so.

Same class names, constructor signatures and method meanings as the reference
(/root/reference/amt_tools/features/*.py); every number is produced by the native library
(integer / float64 frame arithmetic on the host, features by the sm_100a kernels).  Extra,
keyword-only constructor arguments (`device`, `output`) are additions that default to the
drop-in behaviour described in SURVEY.md 8(b):

  process_audio(audio)  audio: 1-D np.ndarray / torch.Tensor (CPU or CUDA)   -> (C, F, T) tensor
                               2-D (B, N) array / tensor                     -> (B, C, F, T) tensor
                               list of 1-D clips (ragged)                    -> list of (C, F, T_b) tensors
  Results are float32 CUDA tensors (T contiguous) on the module's device, or NumPy arrays when the
  module was built with output='numpy'.
"""

import ctypes as C
import json

import numpy as np
import torch

from . import _lib

__all__ = ['FeatureModule', 'WaveformWrapper', 'STFT', 'MelSpec', 'VQT', 'CQT', 'HVQT', 'HCQT', 'SignalPower',
           'FeatureCombo', 'framify_activations']

NOTE_C1_HZ = 440.0 * 2.0 ** ((24 - 69) / 12.0)  # librosa.note_to_hz('C1'), vqt.py:44


class _Plan(object):
    """Owns one amtfeat_plan handle."""

    def __init__(self, cfg, device_index):
        self.handle = C.c_void_p()
        _lib.check(_lib.lib.amtfeat_plan_create(C.byref(cfg), int(device_index), C.byref(self.handle)))
        self.device_index = device_index

    def __del__(self):
        lib = getattr(_lib, 'lib', None)        # None while the interpreter is shutting down
        if lib is not None and getattr(self, 'handle', None) is not None and self.handle.value:
            lib.amtfeat_plan_destroy(self.handle)
            self.handle = C.c_void_p()


def _audio_length(audio):
    return int(audio.shape[-1])


class FeatureModule(object):
    """
    Generic feature extraction module (features/common.py:15).
    """

    _kind = None

    def __init__(self, sample_rate, hop_length, num_channels, decibels=True, *, device=None, output='torch'):
        self.sample_rate = sample_rate
        self.hop_length = hop_length
        self.num_channels = num_channels
        self.decibels = decibels
        if output not in ('torch', 'numpy'):
            raise ValueError("output must be 'torch' or 'numpy'")
        self.output = output
        self.device = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0) \
            if device is None else torch.device(device)
        if self.device.type != 'cuda':
            raise ValueError('amt_tools_b200 modules compute on CUDA devices only (no CPU fallback)')
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._host_plan_ = None
        self._dev_plan_ = None

    # ---- plan management -------------------------------------------------------------------
    def _config(self):
        raise NotImplementedError

    @property
    def _host_plan(self):
        if self._host_plan_ is None:
            self._host_plan_ = _Plan(self._config(), -1)
        return self._host_plan_

    @property
    def _dev_plan(self):
        if self._dev_plan_ is None:
            if not torch.cuda.is_available():
                raise _lib.AmtfeatError('no CUDA device available: amt_tools_b200 has no CPU compute path')
            self._dev_plan_ = _Plan(self._config(), self.device.index)
        return self._dev_plan_

    def describe(self):
        buf = C.create_string_buffer(1 << 18)
        _lib.check(_lib.lib.amtfeat_plan_describe(self._host_plan.handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    def describe_clip(self, num_samples):
        """How one clip of `num_samples` is laid out (frames, ladder levels, exact-ladder tails): tests / docs."""
        buf = C.create_string_buffer(1 << 16)
        _lib.check(_lib.lib.amtfeat_clip_describe(self._host_plan.handle, int(num_samples), buf, len(buf)))
        return json.loads(buf.value.decode())

    # ---- reference API ---------------------------------------------------------------------
    def get_expected_frames(self, audio):
        """features/common.py:41-66 (and overrides)."""
        return int(_lib.lib.amtfeat_expected_frames(self._host_plan.handle, _audio_length(audio)))

    def get_sample_range(self, num_frames):
        """features/common.py:68-97 (and overrides)."""
        lo, hi = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib.amtfeat_sample_range(self._host_plan.handle, int(num_frames), C.byref(lo), C.byref(hi)))
        return np.arange(lo.value, hi.value + 1)

    def get_num_samples_required(self):
        """features/common.py:99-112."""
        return self.get_sample_range(1)[-1]

    @staticmethod
    def divisor_pad(audio, divisor):
        """features/common.py:114-139."""
        pad_amt = divisor - (audio.shape[-1] % divisor)
        if pad_amt > 0 and pad_amt != divisor:
            if isinstance(audio, torch.Tensor):
                audio = torch.nn.functional.pad(audio, (0, int(pad_amt)))
            else:
                audio = np.append(audio, np.zeros(pad_amt).astype(np.float32), axis=-1)
        return audio

    def frame_pad(self, audio):
        """features/common.py:141-166."""
        divisor = self.get_num_samples_required()
        if audio.shape[-1] > divisor:
            divisor = self.hop_length
        return self.divisor_pad(audio, divisor)

    def get_times(self, audio, at_start=False):
        """features/common.py:232-258 (and overrides); float64, bit-exact with librosa.frames_to_time."""
        n = _audio_length(audio)
        T = int(_lib.lib.amtfeat_expected_frames(self._host_plan.handle, n))
        out = np.empty(T, dtype=np.float64)
        _lib.check(_lib.lib.amtfeat_times(self._host_plan.handle, n, int(bool(at_start)),
                                          out.ctypes.data_as(C.POINTER(C.c_double)), T))
        return out

    def get_sample_rate(self):
        return self.sample_rate

    def get_hop_length(self):
        return self.hop_length

    def get_num_channels(self):
        return self.num_channels

    def get_feature_size(self):
        return int(_lib.lib.amtfeat_feature_size(self._host_plan.handle))

    @classmethod
    def features_name(cls):
        """features/common.py:310-321."""
        return cls.__name__

    # alias named by the project brief
    get_feature_tag = features_name

    # ---- compute ---------------------------------------------------------------------------
    def _out_shape(self, n):
        shape = (C.c_int64 * 3)()
        ndim = C.c_int()
        _lib.check(_lib.lib.amtfeat_out_shape(self._host_plan.handle, int(n), shape, C.byref(ndim)))
        return tuple(int(shape[i]) for i in range(ndim.value))

    def _to_device_1d(self, clip):
        if isinstance(clip, torch.Tensor):
            t = clip.detach()
        else:
            t = torch.from_numpy(np.ascontiguousarray(clip, dtype=np.float32))
        if t.dim() != 1:
            raise ValueError('expected mono-channel (1-D) audio, got shape %s' % (tuple(t.shape),))
        return t.to(device=self.device, dtype=torch.float32, non_blocking=True)

    def _pack(self, clips):
        """Concatenate clips into one device buffer with 4-element aligned offsets."""
        lengths = [int(c.shape[-1]) for c in clips]
        if len(clips) == 1:
            t = self._to_device_1d(clips[0]).contiguous()
            if t.data_ptr() % 16 == 0:
                return t, [0], lengths
        offsets, total = [], 0
        for n in lengths:
            offsets.append(total)
            total += (n + 3) // 4 * 4
        buf = torch.empty(max(total, 4), dtype=torch.float32, device=self.device)
        for c, o, n in zip(clips, offsets, lengths):
            if n:
                buf[o:o + n].copy_(self._to_device_1d(c), non_blocking=True)
        return buf, offsets, lengths

    def _batch_layout(self, lengths):
        """Output shapes / offsets / workspace size for a batch with these clip lengths (cached: batches repeat)."""
        key = tuple(lengths)
        cache = self.__dict__.setdefault('_layout_cache', {})
        lay = cache.get(key)
        if lay is None:
            shapes = [self._out_shape(n) for n in lengths]
            sizes = [int(np.prod(s)) for s in shapes]
            out_offsets, total = [], 0
            for sz in sizes:
                out_offsets.append(total)
                total += sz
            n_arr = _lib.i64_array(lengths)
            ws_bytes = int(_lib.lib.amtfeat_workspace_bytes(self._dev_plan.handle, len(lengths), n_arr)) if total else 0
            lay = (shapes, sizes, out_offsets, _lib.i64_array(out_offsets), n_arr, ws_bytes, total)
            if len(cache) > 64:
                cache.clear()
            cache[key] = lay
        return lay

    def _launch(self, buf, offsets, lengths, raw=False):
        """Enqueue the native path on the current stream; returns the flat output buffer and the batch layout.
        `raw`: without the dB epilogue (amtfeat_process_raw; chunks of a long track, longtrack.py)."""
        plan = self._dev_plan
        lay = self._batch_layout(lengths)
        shapes, sizes, out_offsets, out_arr, n_arr, ws_bytes, total = lay
        with torch.cuda.device(self.device):
            out = torch.empty(max(total, 1), dtype=torch.float32, device=self.device)
            if total:
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.device)
                stream = torch.cuda.current_stream(self.device).cuda_stream
                in_arr = _lib.i64_array(offsets) if isinstance(offsets, (list, tuple)) else offsets
                fn = _lib.lib.amtfeat_process_raw if raw else _lib.lib.amtfeat_process
                _lib.check(fn(plan.handle, buf.data_ptr(), in_arr, n_arr, out_arr,
                              len(lengths), out.data_ptr(), ws.data_ptr(), ws_bytes, stream))
                # the caching allocator keeps `ws` / `buf` alive for this stream's pending work
        return out, lay

    def _run(self, buf, offsets, lengths):
        """Launch and return one tensor per clip."""
        out, (shapes, sizes, out_offsets, _, _, _, _) = self._launch(buf, offsets, lengths)
        return [out[o:o + sz].view(shape) for o, sz, shape in zip(out_offsets, sizes, shapes)]

    def _finish(self, t):
        return t.cpu().numpy() if self.output == 'numpy' else t

    def process_audio(self, audio):
        """
        Features for a piece (or batch) of audio -- stft.py:42, mel.py:40, vqt.py:167, hvqt.py:107,
        power.py:31, waveform.py:121.
        """
        if isinstance(audio, (list, tuple)):
            buf, offsets, lengths = self._pack(list(audio))
            return [self._finish(t) for t in self._run(buf, offsets, lengths)]
        if audio.ndim == 2:
            B, N = int(audio.shape[0]), int(audio.shape[1])
            if isinstance(audio, torch.Tensor) and audio.is_cuda and audio.dtype == torch.float32 \
                    and audio.is_contiguous() and N % 4 == 0 and audio.data_ptr() % 16 == 0:
                offs = self.__dict__.setdefault('_uniform_offsets', {})
                if (B, N) not in offs:
                    offs[(B, N)] = _lib.i64_array([b * N for b in range(B)])
                out, lay = self._launch(audio.reshape(-1), offs[(B, N)], [N] * B)
                shape, per = lay[0][0], lay[1][0]
                if per:     # equal-length clips are written back to back: one (B, ...) tensor, no copy
                    return self._finish(out[:B * per].view((B,) + shape))
                outs = [out[:0].view(shape) for _ in range(B)]
            else:
                buf, offsets, lengths = self._pack([audio[b] for b in range(B)])
                outs = self._run(buf, offsets, lengths)
            # clips of equal length were written back to back: expose them as one (B, ...) tensor
            base = outs[0]
            full = torch.as_strided(base, (B,) + tuple(base.shape), (base.numel(),) + tuple(base.stride()),
                                    base.storage_offset()) if base.numel() else torch.stack(outs)
            return self._finish(full)
        buf, offsets, lengths = self._pack([audio])
        return self._finish(self._run(buf, offsets, lengths)[0])


class WaveformWrapper(FeatureModule):
    """Audio framing wrapper (features/waveform.py:14)."""

    _kind = _lib.WAVEFORM

    def __init__(self, sample_rate=44100, hop_length=512, decibels=False, win_length=None, center=True, **kw):
        super().__init__(sample_rate=sample_rate, hop_length=hop_length, num_channels=1, decibels=decibels, **kw)
        if win_length is None:
            win_length = self.hop_length
        self.win_length = win_length
        self.center = center

    def _config(self):
        return _lib.Config(kind=self._kind, hop_length=int(self.hop_length), sample_rate=float(self.sample_rate),
                           decibels=int(bool(self.decibels)), center=int(bool(self.center)),
                           win_length=int(self.win_length), n_fft=int(getattr(self, 'n_fft', 0)),
                           n_mels=int(getattr(self, 'n_mels', 0)), htk=int(bool(getattr(self, 'htk', False))))

    def center_pad(self, audio):
        """features/waveform.py:98-119."""
        pad = int(self.win_length // 2)
        if isinstance(audio, torch.Tensor):
            return torch.nn.functional.pad(audio, (pad, pad))
        return np.pad(audio, [(pad, pad)], mode='constant')


class STFT(WaveformWrapper):
    """Magnitude spectrogram (features/stft.py:11)."""

    _kind = _lib.STFT

    def __init__(self, sample_rate=16000, hop_length=512, decibels=True, win_length=None, center=True, n_fft=2048,
                 **kw):
        self.n_fft = n_fft
        if win_length is None:
            win_length = self.n_fft
        super().__init__(sample_rate=sample_rate, hop_length=hop_length, decibels=decibels, win_length=win_length,
                         center=center, **kw)


class MelSpec(STFT):
    """Mel spectrogram (features/mel.py:11)."""

    _kind = _lib.MEL

    def __init__(self, sample_rate=16000, hop_length=512, decibels=True, n_mels=229, n_fft=2048, win_length=None,
                 center=True, htk=False, **kw):
        super().__init__(sample_rate=sample_rate, hop_length=hop_length, decibels=decibels, win_length=win_length,
                         center=center, n_fft=n_fft, **kw)
        self.n_mels = n_mels
        self.htk = htk


class SignalPower(WaveformWrapper):
    """Frame-level signal power (features/power.py:12)."""

    _kind = _lib.POWER

    def __init__(self, sample_rate=44100, hop_length=512, decibels=True, win_length=None, center=True, **kw):
        super().__init__(sample_rate=sample_rate, hop_length=hop_length, decibels=decibels, win_length=win_length,
                         center=center, **kw)


class VQT(FeatureModule):
    """Variable-Q transform (features/vqt.py:17)."""

    _kind = _lib.VQT

    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, n_bins=84, bins_per_octave=12,
                 gamma=None, **kw):
        super().__init__(sample_rate, hop_length, 1, decibels, **kw)
        if fmin is None:
            fmin = NOTE_C1_HZ
        self.fmin = fmin
        self.n_bins = n_bins
        self.bins_per_octave = bins_per_octave
        self.window = 'hann'
        self.alpha = 2.0 ** (1.0 / self.bins_per_octave) - 1          # vqt.py:49
        if gamma is None:
            gamma = 24.7 * self.alpha / 0.108                          # vqt.py:52-58
        self.gamma = gamma
        self.n_octs = int(np.ceil(float(self.n_bins) / self.bins_per_octave))

    def _harmonics(self):
        return [1.0]

    def _config(self):
        cfg = _lib.Config(kind=self._kind, hop_length=int(self.hop_length), sample_rate=float(self.sample_rate),
                          decibels=int(bool(self.decibels)), n_bins=int(self.n_bins),
                          bins_per_octave=int(self.bins_per_octave), fmin=float(self.fmin), gamma=float(self.gamma))
        hs = self._harmonics()
        if len(hs) > _lib.MAX_HARMONICS:
            raise ValueError('at most %d harmonics are supported' % _lib.MAX_HARMONICS)
        cfg.n_harmonics = len(hs)
        for i, h in enumerate(hs):
            cfg.harmonics[i] = float(h)
        return cfg

    def get_early_ds_count(self):
        """features/vqt.py:64-100."""
        return int(_lib.lib.amtfeat_early_ds_count(self._host_plan.handle, 0))


class CQT(VQT):
    """Constant-Q transform = VQT with gamma = 0 (features/cqt.py:7)."""

    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, n_bins=84, bins_per_octave=12,
                 **kw):
        super().__init__(sample_rate, hop_length, decibels, fmin, n_bins, bins_per_octave, gamma=0, **kw)


class HVQT(FeatureModule):
    """
    Harmonic VQT (features/hvqt.py:12).  The reference runs one independent VQT per harmonic; here all
    harmonics share one decimation ladder and one FFT per (ladder level, n_fft) inside a single plan.
    `self.modules` still holds one VQT per harmonic for API compatibility (each is usable on its own).
    """

    _kind = _lib.HVQT

    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, harmonics=None, n_bins=84,
                 bins_per_octave=12, gamma=None, **kw):
        if fmin is None:
            fmin = NOTE_C1_HZ
        self.fmin = fmin
        if harmonics is None:
            harmonics = [0.5, 1, 2, 3, 4, 5]
        harmonics.sort()  # in place, like hvqt.py:40
        self.harmonics = harmonics
        super().__init__(sample_rate, hop_length, len(self.harmonics), decibels, **kw)
        self.n_bins = n_bins
        self.bins_per_octave = bins_per_octave
        alpha = 2.0 ** (1.0 / bins_per_octave) - 1
        self.gamma = 24.7 * alpha / 0.108 if gamma is None else gamma
        self.modules = [VQT(sample_rate=sample_rate, hop_length=hop_length, decibels=decibels, fmin=h * fmin,
                            n_bins=n_bins, bins_per_octave=bins_per_octave, gamma=gamma, **kw)
                        for h in self.harmonics]

    def _config(self):
        cfg = _lib.Config(kind=self._kind, hop_length=int(self.hop_length), sample_rate=float(self.sample_rate),
                          decibels=int(bool(self.decibels)), n_bins=int(self.n_bins),
                          bins_per_octave=int(self.bins_per_octave), fmin=float(self.fmin), gamma=float(self.gamma))
        if len(self.harmonics) > _lib.MAX_HARMONICS:
            raise ValueError('at most %d harmonics are supported' % _lib.MAX_HARMONICS)
        cfg.n_harmonics = len(self.harmonics)
        for i, h in enumerate(self.harmonics):
            cfg.harmonics[i] = float(h)
        return cfg

    def to_decibels(self, feats):
        """features/hvqt.py:135-146 (dB is applied per harmonic inside the kernels)."""
        return NotImplementedError


class HCQT(HVQT):
    """Harmonic CQT = HVQT with gamma = 0 (features/hcqt.py:7)."""

    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, harmonics=None, n_bins=84,
                 bins_per_octave=12, **kw):
        super().__init__(sample_rate, hop_length, decibels, fmin, harmonics, n_bins, bins_per_octave, gamma=0, **kw)


class FeatureCombo(FeatureModule):
    """
    Combination of feature modules (features/combo.py:14).  Like the reference it does not call the base
    constructor.  `process_audio` concatenates along the channel axis exactly as the reference does (and
    fails the same way when the shapes do not line up, combo.py:118-120); `process_audio_list` is the
    addition that returns the per-module results, uploading the audio to the device once.
    """

    def __init__(self, modules):
        self.modules = modules

    def get_expected_frames(self, audio):
        num_frames = [module.get_expected_frames(audio) for module in self.modules]
        assert len(set(num_frames)) == 1
        return num_frames[0]

    def get_sample_range(self, num_frames):
        sample_range = None
        for module in self.modules:
            r = module.get_sample_range(num_frames)
            sample_range = r if sample_range is None else np.intersect1d(sample_range, r)
        return sample_range

    def _shared_upload(self, audio):
        dev = self.modules[0].device
        if isinstance(audio, (list, tuple)) or not all(m.device == dev for m in self.modules):
            return audio
        if isinstance(audio, torch.Tensor):
            return audio.to(device=dev, dtype=torch.float32)
        return torch.from_numpy(np.ascontiguousarray(audio, dtype=np.float32)).to(dev)

    def process_audio_list(self, audio):
        audio = self._shared_upload(audio)
        dev = self.modules[0].device
        if len(self.modules) < 2 or not torch.cuda.is_available() or not all(m.device == dev for m in self.modules) \
                or any(m.output == 'numpy' for m in self.modules):
            return [module.process_audio(audio) for module in self.modules]
        # The modules are independent: each runs on its own stream (forked from / joined to the caller's stream), so a
        # compute-bound kernel of one module runs underneath the HBM-bound dB epilogue or the short ladder levels of another.
        streams = self.__dict__.setdefault('_streams', [torch.cuda.Stream(dev) for _ in self.modules])
        cur = torch.cuda.current_stream(dev)
        feats = []
        for module, s in zip(self.modules, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                f = module.process_audio(audio)
            for t in (f if isinstance(f, (list, tuple)) else [f]):
                if isinstance(t, torch.Tensor):
                    t.record_stream(cur)   # produced on a side stream, consumed on the caller's
            feats.append(f)
        for s in streams:
            cur.wait_stream(s)
        return feats

    def process_audio(self, audio):
        feats = [f for f in self.process_audio_list(audio) if f is not None]
        if len(feats) == 0:
            return None
        if any(isinstance(f, np.ndarray) for f in feats):
            return np.concatenate([np.asarray(f) for f in feats], axis=0)
        try:
            return torch.cat(feats, dim=0)
        except RuntimeError as e:  # np.concatenate raises ValueError in the reference
            raise ValueError(str(e))

    def get_times(self, audio):
        return [module.get_times(audio) for module in self.modules][0]

    def get_sample_rate(self):
        sample_rate = [module.get_sample_rate() for module in self.modules]
        assert len(set(sample_rate)) == 1
        return sample_rate[0]

    def get_hop_length(self):
        hop_length = [module.get_hop_length() for module in self.modules]
        assert len(set(hop_length)) == 1
        return hop_length[0]

    def get_num_channels(self):
        return sum([module.get_num_channels() for module in self.modules])

    def get_feature_size(self):
        return NotImplementedError


def framify_activations(activations, win_length, hop_length=1, pad=True):
    """
    Device-side `tools.framify_activations` (amt_tools/tools/utils.py:2922-2984): chunk a CUDA tensor (..., T) into
    overlapping frames (..., num_hops, win_length) without leaving the device (TabCNN.pre_proc, models/tabcnn.py:123-127).
    """
    if not (isinstance(activations, torch.Tensor) and activations.is_cuda):
        raise ValueError('framify_activations expects a CUDA tensor (there is no CPU compute path)')
    x = activations.to(torch.float32).contiguous()
    T = int(x.shape[-1])
    rows = int(x.numel() // max(T, 1)) if T else int(np.prod(x.shape[:-1]))
    hops = int(_lib.lib.amtfeat_framify_hops(T, int(win_length), int(hop_length), int(bool(pad))))
    if hops < 0:
        raise ValueError('invalid framify arguments')
    with torch.cuda.device(x.device):
        out = torch.empty(tuple(x.shape[:-1]) + (hops, int(win_length)), dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib.amtfeat_framify(x.data_ptr(), rows, T, int(win_length), int(hop_length), int(bool(pad)),
                                            out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream))
    return out
