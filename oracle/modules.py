"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product (amt_tools_b200/).

CPU restatement of the reference's FeatureModule wrapper logic (frame-count / sample-range /
time arithmetic, padding, dB post-processing, harmonic stacking), layered over
oracle.librosa_stages.  Each method cites the reference lines it follows
(paths under /root/reference/amt_tools/features/).

`dtype=np.float32` mirrors the reference's working precision (float32 audio in, complex64
spectra); `dtype=np.float64` is the ground truth the GPU tolerances are measured against.
"""

import numpy as np

from . import librosa_stages as ls


class OFeatureModule(object):
    # common.py:20-39
    def __init__(self, sample_rate, hop_length, num_channels, decibels=True, dtype=np.float64):
        self.sample_rate, self.hop_length = sample_rate, hop_length
        self.num_channels, self.decibels = num_channels, decibels
        self.dtype = dtype

    # common.py:41-66
    def get_expected_frames(self, audio):
        return 0 if audio.shape[-1] == 0 else 1 + len(audio) // self.hop_length

    # common.py:68-97
    def get_sample_range(self, num_frames):
        if num_frames <= 0:
            return np.array([0])
        hi = num_frames * self.hop_length - 1
        lo = max(1, hi - self.hop_length + 1)
        return np.arange(lo, hi + 1)

    # common.py:99-112
    def get_num_samples_required(self):
        return self.get_sample_range(1)[-1]

    # common.py:114-139
    @staticmethod
    def divisor_pad(audio, divisor):
        pad = divisor - (audio.shape[-1] % divisor)
        if 0 < pad != divisor:
            audio = np.append(audio, np.zeros(pad, dtype=np.float32), axis=-1)
        return audio

    # common.py:141-166
    def frame_pad(self, audio):
        divisor = self.get_num_samples_required()
        if audio.shape[-1] > divisor:
            divisor = self.hop_length
        return self.divisor_pad(audio, divisor)

    # common.py:181-201
    def to_decibels(self, feats):
        return ls.amplitude_to_db(feats, ref=np.max)

    # common.py:203-230
    def post_proc(self, feats):
        if self.decibels:
            feats = self.to_decibels(feats)
            feats = feats / 80
            feats = feats + 1
        return np.expand_dims(feats, axis=0)

    # common.py:232-258
    def get_times(self, audio):
        return ls.frames_to_time(np.arange(self.get_expected_frames(audio)), self.sample_rate, self.hop_length)


class OWaveformWrapper(OFeatureModule):
    # waveform.py:18-41
    def __init__(self, sample_rate=44100, hop_length=512, decibels=False, win_length=None, center=True,
                 dtype=np.float64):
        super().__init__(sample_rate, hop_length, 1, decibels, dtype)
        self.win_length = hop_length if win_length is None else win_length
        self.center = center

    # waveform.py:43-66
    def get_expected_frames(self, audio):
        if self.center or audio.shape[-1] == 0:
            return super().get_expected_frames(audio)
        return 1 + ((max(0, audio.shape[-1] - self.win_length) - 1) // self.hop_length + 1)

    # waveform.py:68-96
    def get_sample_range(self, num_frames):
        if self.center or num_frames == 0:
            return super().get_sample_range(num_frames)
        if num_frames == 1:
            return np.arange(1, self.win_length + 1)
        return np.arange(1, self.hop_length + 1) + self.get_num_samples_required() + (num_frames - 2) * self.hop_length

    # waveform.py:98-119
    def center_pad(self, audio):
        return np.pad(audio, int(self.win_length // 2), mode='constant')

    # waveform.py:121-153
    def process_audio(self, audio):
        if audio.shape[-1] == 0:
            return np.zeros((self.win_length, 0))
        audio = self.center_pad(audio) if self.center else self.frame_pad(audio)
        return ls.frame(audio, self.win_length, self.hop_length)

    # waveform.py:155-185
    def get_times(self, audio, at_start=False):
        times = super().get_times(audio)
        if self.center and at_start:
            times -= (self.win_length // 2) / self.sample_rate
        elif not self.center and not at_start:
            times += (self.win_length // 2) / self.sample_rate
        return times

    def get_feature_size(self):
        return self.win_length


class OSTFT(OWaveformWrapper):
    # stft.py:15-40
    def __init__(self, sample_rate=16000, hop_length=512, decibels=True, win_length=None, center=True,
                 n_fft=2048, dtype=np.float64):
        self.n_fft = n_fft
        super().__init__(sample_rate, hop_length, decibels, n_fft if win_length is None else win_length,
                         center, dtype)

    # stft.py:42-77
    def process_audio(self, audio):
        if audio.shape[-1] == 0:
            return np.zeros((1, self.n_fft, 0))
        if not self.center:
            audio = self.frame_pad(audio)
        spec = np.abs(ls.stft(audio, n_fft=self.n_fft, hop_length=self.hop_length,
                              win_length=self.win_length, center=self.center, dtype=self.dtype))
        return OFeatureModule.post_proc(self, spec)

    def get_feature_size(self):
        return self.n_fft // 2 + 1


class OMelSpec(OSTFT):
    # mel.py:15-38
    def __init__(self, sample_rate=16000, hop_length=512, decibels=True, n_mels=229, n_fft=2048,
                 win_length=None, center=True, htk=False, dtype=np.float64):
        super().__init__(sample_rate, hop_length, decibels, win_length, center, n_fft, dtype)
        self.n_mels, self.htk = n_mels, htk

    # mel.py:40-76
    def process_audio(self, audio):
        if audio.shape[-1] == 0:
            return np.zeros((1, self.n_mels, 0))
        if not self.center:
            audio = self.frame_pad(audio)
        mel = ls.melspectrogram(audio, self.sample_rate, self.n_fft, self.hop_length, self.win_length,
                                self.center, self.htk, self.n_mels, dtype=self.dtype)
        return OFeatureModule.post_proc(self, mel)

    # mel.py:78-96
    def to_decibels(self, feats):
        return ls.power_to_db(feats, ref=np.max)

    def get_feature_size(self):
        return self.n_mels


class OSignalPower(OWaveformWrapper):
    # power.py:16-29
    def __init__(self, sample_rate=44100, hop_length=512, decibels=True, win_length=None, center=True,
                 dtype=np.float64):
        super().__init__(sample_rate, hop_length, decibels, win_length, center, dtype)

    # power.py:31-57
    def process_audio(self, audio):
        frames = super().process_audio(np.asarray(audio, dtype=self.dtype))
        powers = np.sum(frames ** 2, axis=-2) / self.win_length
        if self.decibels:
            powers = ls.amplitude_to_db(powers, ref=np.max)
        return powers

    def get_feature_size(self):
        return 1


class OVQT(OFeatureModule):
    # vqt.py:21-62
    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, n_bins=84,
                 bins_per_octave=12, gamma=None, dtype=np.float64, basis_cache=None):
        super().__init__(sample_rate, hop_length, 1, decibels, dtype)
        self.fmin = ls.NOTE_C1_HZ if fmin is None else fmin
        self.n_bins, self.bins_per_octave = n_bins, bins_per_octave
        self.alpha = 2.0 ** (1.0 / bins_per_octave) - 1
        self.gamma = 24.7 * self.alpha / 0.108 if gamma is None else gamma
        self.n_octs = int(np.ceil(float(n_bins) / bins_per_octave))
        self.basis_cache = basis_cache

    # vqt.py:64-100
    def get_early_ds_count(self):
        fmax = np.max(ls.cqt_frequencies(self.n_bins, self.fmin, self.bins_per_octave))
        cQ = 1.0 / (2.0 ** (1. / self.bins_per_octave) - 1)
        cutoff = fmax * (1 + 0.5 * ls.HANN_BANDWIDTH / cQ) + 0.5 * self.gamma
        return ls.early_downsample_count(self.sample_rate / 2.0, cutoff, self.hop_length, self.n_octs)

    # vqt.py:102-134
    def get_expected_frames(self, audio):
        eds = self.get_early_ds_count()
        k = np.arange(eds, eds + self.n_octs)
        sig_lens = np.ceil(len(audio) / (2 ** k))
        hop_lens = self.hop_length // (2 ** k)
        return int(min(sig_lens // hop_lens + 1))

    # vqt.py:136-165
    def get_sample_range(self, num_frames):
        f = 2 ** self.get_early_ds_count()
        hi = ((num_frames * self.hop_length // f) - 1) * f
        lo = max(1, hi - self.hop_length + 1)
        return np.arange(lo, hi + 1)

    # vqt.py:167-195
    def process_audio(self, audio):
        v = np.abs(ls.vqt(audio, sr=self.sample_rate, hop_length=self.hop_length, fmin=self.fmin,
                          n_bins=self.n_bins, bins_per_octave=self.bins_per_octave, gamma=self.gamma,
                          dtype=self.dtype, basis_cache=self.basis_cache))
        return self.post_proc(v)

    # vqt.py:197-227 (intended behaviour: (L_fmin // 2) / sr with the module's own alpha / gamma)
    def get_times(self, audio, at_start=False):
        times = super().get_times(audio)
        if at_start:
            longest, _ = ls.wavelet_lengths(self.fmin, self.sample_rate, self.gamma, self.alpha)
            times -= (longest[0] // 2) / self.sample_rate
        return times

    def get_feature_size(self):
        return self.n_bins


class OCQT(OVQT):
    # cqt.py:12-22
    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, n_bins=84,
                 bins_per_octave=12, dtype=np.float64, basis_cache=None):
        super().__init__(sample_rate, hop_length, decibels, fmin, n_bins, bins_per_octave, gamma=0,
                         dtype=dtype, basis_cache=basis_cache)


class OHVQT(OFeatureModule):
    # hvqt.py:16-58
    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, harmonics=None,
                 n_bins=84, bins_per_octave=12, gamma=None, dtype=np.float64, basis_cache=None):
        self.fmin = ls.NOTE_C1_HZ if fmin is None else fmin
        if harmonics is None:
            harmonics = [0.5, 1, 2, 3, 4, 5]
        harmonics.sort()
        self.harmonics = harmonics
        super().__init__(sample_rate, hop_length, len(harmonics), decibels, dtype)
        self.modules = [OVQT(sample_rate, hop_length, decibels, h * self.fmin, n_bins, bins_per_octave, gamma,
                             dtype=dtype, basis_cache=basis_cache) for h in harmonics]

    # hvqt.py:60-83
    def get_expected_frames(self, audio):
        return min(m.get_expected_frames(audio) for m in self.modules)

    # hvqt.py:85-105
    def get_sample_range(self, num_frames):
        return self.modules[-1].get_sample_range(num_frames)

    # hvqt.py:107-133
    def process_audio(self, audio):
        T = self.get_expected_frames(audio)
        return np.concatenate([m.process_audio(audio)[..., :T] for m in self.modules], axis=0)

    # hvqt.py:148-168
    def get_times(self, audio, at_start=False):
        return self.modules[0].get_times(audio, at_start)[:self.get_expected_frames(audio)]

    def get_feature_size(self):
        return self.modules[0].get_feature_size()


class OHCQT(OHVQT):
    # hcqt.py:11-21
    def __init__(self, sample_rate=22050, hop_length=512, decibels=True, fmin=None, harmonics=None,
                 n_bins=84, bins_per_octave=12, dtype=np.float64, basis_cache=None):
        super().__init__(sample_rate, hop_length, decibels, fmin, harmonics, n_bins, bins_per_octave, gamma=0,
                         dtype=dtype, basis_cache=basis_cache)


class OFeatureCombo(object):
    # combo.py:18-28 (does not call the base constructor)
    def __init__(self, modules):
        self.modules = modules

    # combo.py:30-55
    def get_expected_frames(self, audio):
        counts = [m.get_expected_frames(audio) for m in self.modules]
        assert len(set(counts)) == 1
        return counts[0]

    # combo.py:57-84
    def get_sample_range(self, num_frames):
        rng = None
        for m in self.modules:
            r = m.get_sample_range(num_frames)
            rng = r if rng is None else np.intersect1d(rng, r)
        return rng

    # combo.py:86-122
    def process_audio(self, audio):
        feats = [f for f in (m.process_audio(audio) for m in self.modules) if f is not None]
        return np.concatenate(feats, axis=0) if feats else None

    def process_audio_list(self, audio):
        return [m.process_audio(audio) for m in self.modules]

    # combo.py:124-150
    def get_times(self, audio):
        return self.modules[0].get_times(audio)

    # combo.py:192-204
    def get_num_channels(self):
        return sum(m.num_channels for m in self.modules)
