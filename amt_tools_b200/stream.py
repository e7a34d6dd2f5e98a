"""
Streaming wrappers over a FeatureModule -- the step *after* the hot path.

Mirrors `amt_tools.features.stream.FeatureStream` / `AudioStream` (/root/reference/amt_tools/features/stream.py:33-254,
637-779): same constructor arguments, same methods, same buffer semantics, same result dictionary
(`{'features': (1, C, F, T_buffered), 'times': (1, 1)}`).  The reference recomputes `process_audio` on a slice of
`get_num_samples_required()` samples for every hop (stream.py:746-755), one librosa call per frame.  Here the slices of
the next `lookahead` hops are independent clips of one ragged batch, so they go through ONE launch of the module's
kernels (per-clip dB reference maximum, exactly the per-slice `ref=np.max` the reference applies) and are then handed out
hop by hop.  `lookahead=1` is the reference's one-slice-per-call behaviour.

`MicrophoneStream` / `AudioFileStream` need audio hardware or a decoder and stay out of scope.
"""

import time

import numpy as np
import torch

KEY_FEATS = 'features'   # amt_tools/tools/constants.py:50
KEY_TIMES = 'times'      # amt_tools/tools/constants.py:56
MIC_LAG_TOL = 0.250      # seconds (stream.py:30)


def _current_time(decimals=3):
    return round(time.time(), decimals)


class FeatureStream(object):
    """Generic feature streaming wrapper (stream.py:33-254)."""

    def __init__(self, module, frame_buffer_size=1):
        self.module = module
        self.frame_buffer = None
        self.frame_buffer_size = frame_buffer_size
        self.start_time = None

    def reset_stream(self):
        self.stop_streaming()
        self.frame_buffer = list()

    def start_streaming(self):
        self.start_time = _current_time()

    def stop_streaming(self):
        self.start_time = None

    def extract_frame_features(self):
        return NotImplementedError   # the reference returns (not raises) it: stream.py:96

    def query_active(self):
        return self.start_time is not None

    def query_finished(self):
        return NotImplementedError

    def buffer_new_frame(self, frame=None):
        if frame is None:
            frame = self.extract_frame_features()
        if self.query_frame_buffer_full():
            # make room for exactly one more frame
            start_idx = len(self.frame_buffer) - self.frame_buffer_size + 1
            self.frame_buffer = self.frame_buffer[start_idx:]
        self.frame_buffer += [frame]
        return self.get_buffered_frames()

    def _empty_frame(self):
        shape = (self.module.get_num_channels(), self.module.get_feature_size(), 1)
        if getattr(self.module, 'output', 'torch') == 'numpy' or getattr(self.module, 'device', None) is None:
            return np.zeros(shape, dtype=np.float32)
        return torch.zeros(shape, dtype=torch.float32, device=self.module.device)

    def buffer_empty_frame(self):
        return self.buffer_new_frame(self._empty_frame())

    def prime_frame_buffer(self, amount):
        for _ in range(amount):
            self.buffer_empty_frame()

    def query_frame_buffer_full(self):
        return len(self.frame_buffer) >= self.frame_buffer_size

    def get_buffered_frames(self):
        times = np.array([self.get_elapsed_time()])
        if any(isinstance(f, torch.Tensor) for f in self.frame_buffer):
            dev = next(f.device for f in self.frame_buffer if isinstance(f, torch.Tensor))
            feats = torch.cat([f if isinstance(f, torch.Tensor) else torch.from_numpy(f).to(dev) for f in self.frame_buffer], dim=-1)
            return {KEY_FEATS: feats.unsqueeze(0), KEY_TIMES: times[None]}
        feats = np.concatenate(self.frame_buffer, axis=-1)
        return {KEY_FEATS: feats[None], KEY_TIMES: times[None]}   # tools.dict_unsqueeze: a leading batch axis on every entry

    def get_elapsed_time(self, decimals=3):
        elapsed = 0
        if self.start_time is not None:
            elapsed = round(_current_time(decimals) - self.start_time, decimals)
        return elapsed


class AudioStream(FeatureStream):
    """
    Streams features of an in-memory signal hop by hop (stream.py:637-779).

    lookahead : int (keyword only, new)
      Hops whose slices are processed together in one batched launch.  1 = one slice per call (the reference).
    """

    def __init__(self, module, frame_buffer_size=1, audio=None, real_time=False, playback=False, suppress_warnings=True,
                 *, lookahead=1):
        FeatureStream.__init__(self, module, frame_buffer_size)
        if playback:
            raise ValueError('playback needs audio hardware (sounddevice): out of scope of the device feature path')
        self.audio = None
        self.current_sample = None
        self.playback = playback
        self.real_time = real_time
        self.suppress_warnings = suppress_warnings
        self.lookahead = max(1, int(lookahead))
        self._ready = []          # features of the next hops, already computed
        self.reset_stream(audio)

    def reset_stream(self, audio=None):
        super().reset_stream()
        self.current_sample = 0
        self._ready = []
        if audio is not None:
            self.audio = audio

    def query_finished(self):
        finished = True
        if self.audio is not None:
            finished = self.current_sample > len(self.audio)   # stream.py:777 (a last, empty slice is still served)
        return finished

    def _refill(self):
        need = self.module.get_num_samples_required()
        hop = self.module.get_hop_length()
        starts, s = [], self.current_sample
        while len(starts) < self.lookahead and s <= len(self.audio):
            starts.append(s)
            s += hop
        slices = [self.audio[..., a:a + need] for a in starts]
        if len(slices) == 1:
            self._ready = [self.module.process_audio(slices[0])]
        else:
            self._ready = list(self.module.process_audio(slices))   # one ragged batch, one launch per kernel

    def extract_frame_features(self):
        features = None
        if self.query_active() and not self.query_finished():
            sample_time = (self.current_sample + self.module.get_num_samples_required()) / self.module.sample_rate
            if self.real_time:
                if not self.suppress_warnings:
                    lag = self.get_elapsed_time() - sample_time
                    if lag > MIC_LAG_TOL:
                        import warnings
                        warnings.warn('Processing might be too slow. Currently out of sync by %s seconds.' % lag,
                                      category=RuntimeWarning)
                while self.get_elapsed_time() < sample_time:   # wait until the slice would have been recorded
                    continue
            if not self._ready:
                self._refill()
            features = self._ready.pop(0)
            self.current_sample += self.module.get_hop_length()
        return features
