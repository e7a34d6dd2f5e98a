"""
Streaming wrappers over a FeatureModule -- the step *after* the hot path.

Mirrors `amt_tools.features.stream.FeatureStream` / `AudioStream` (/root/reference/amt_tools/features/stream.py:33-254,
637-779): same constructor arguments, same methods, same buffer semantics, same result dictionary
(`{'features': (1, C, F, T_buffered), 'times': (1, 1)}`).  The reference recomputes `process_audio` on a slice of
`get_num_samples_required()` samples for every hop (stream.py:746-755), one librosa call per frame.  Here the slices of
the next `lookahead` hops are independent clips of one ragged batch, so they go through ONE launch of the module's
kernels (per-clip dB reference maximum, exactly the per-slice `ref=np.max` the reference applies) and are then handed out
hop by hop.  `lookahead=1` is the reference's one-slice-per-call behaviour.  The frame buffer is a preallocated ring on the
module's device (`_FrameRing`): one column write per hop and one gather per buffered view, where the reference concatenates
its whole list of frames on every hop.

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


class _FrameRing(object):
    """
    The last `size` frames in arrival order.  The reference keeps a Python list and concatenates all of it on every hop
    (stream.py:118-142, 233-254): `size` small copies per hop.  Here the one-frame-wide features a stream produces live in ONE
    preallocated (C, F, size) block on the device (or in numpy for host modules): a hop is one column write and the buffered
    view is one gather of at most two slices.  Anything that does not fit the block (the zero-width frame of a final, empty
    slice; frames of another type or shape) drops the ring into list form, which behaves like the reference's list.
    Reads like a list: len(), iteration oldest first, == [].
    """

    def __init__(self):
        self.block, self.head, self.count, self.loose = None, 0, 0, None

    def __len__(self):
        return len(self.loose) if self.loose is not None else self.count

    def __iter__(self):
        if self.loose is not None:
            return iter(self.loose)
        cap = self.block.shape[-1] if self.block is not None else 1
        return iter([self.block[..., (self.head + i) % cap:(self.head + i) % cap + 1] for i in range(self.count)])

    def __eq__(self, other):
        mine = list(self)
        return isinstance(other, (list, _FrameRing)) and len(mine) == len(other) and all(
            np.array_equal(np.asarray(torch.as_tensor(a).cpu()), np.asarray(torch.as_tensor(b).cpu())) for a, b in zip(mine, other))

    def _fits(self, frame, size):
        if self.loose is not None or not isinstance(frame, (np.ndarray, torch.Tensor)) or frame.ndim != 3 or frame.shape[-1] != 1:
            return False
        if self.block is None or self.block.shape[-1] != size:
            return self.count == 0 or (type(frame) is type(self.block) and tuple(frame.shape[:2]) == tuple(self.block.shape[:2]))
        same_place = not isinstance(frame, torch.Tensor) or (frame.device == self.block.device and frame.dtype == self.block.dtype)
        return type(frame) is type(self.block) and tuple(frame.shape[:2]) == tuple(self.block.shape[:2]) and same_place

    def _rebuild(self, frame, size):
        """(Re)allocate the block for `size` columns like `frame`, keeping the newest frames held."""
        keep = list(self)[-(size - 1):] if size > 1 else []
        shape = tuple(frame.shape[:2]) + (size,)
        self.block = torch.empty(shape, dtype=frame.dtype, device=frame.device) if isinstance(frame, torch.Tensor) else np.empty(shape, dtype=frame.dtype)
        self.head, self.count = 0, 0
        for f in keep:
            self.block[..., self.count:self.count + 1] = f
            self.count += 1

    def push(self, frame, size):
        size = max(1, int(size))
        if self._fits(frame, size):
            if self.block is None or self.block.shape[-1] != size:
                self._rebuild(frame, size)
            if self.count == size:                       # full: the oldest column is the one overwritten
                slot, self.head = self.head, (self.head + 1) % size
            else:
                slot, self.count = (self.head + self.count) % size, self.count + 1
            self.block[..., slot:slot + 1] = frame
            return
        if self.loose is None:
            self.loose = [f.clone() if isinstance(f, torch.Tensor) else np.array(f) for f in self]
            self.block, self.head, self.count = None, 0, 0
        if len(self.loose) >= size:                      # make room for exactly one more frame (stream.py:132-136)
            self.loose = self.loose[len(self.loose) - size + 1:]
        self.loose.append(frame)

    def stacked(self):
        """All frames side by side, oldest first, as a NEW array (a caller may keep it across hops)."""
        if self.loose is not None:
            if any(isinstance(f, torch.Tensor) for f in self.loose):
                dev = next(f.device for f in self.loose if isinstance(f, torch.Tensor))
                return torch.cat([f if isinstance(f, torch.Tensor) else torch.from_numpy(f).to(dev) for f in self.loose], dim=-1)
            return np.concatenate(self.loose, axis=-1)
        if self.block is None:
            return np.concatenate([], axis=-1)           # raises like the reference's np.concatenate of an empty list
        size = self.block.shape[-1]
        first = self.block[..., self.head:min(size, self.head + self.count)]
        rest = self.block[..., :max(0, self.head + self.count - size)]
        if isinstance(self.block, torch.Tensor):
            return torch.cat([first, rest], dim=-1)
        return np.concatenate([first, rest], axis=-1)


class FeatureStream(object):
    """Generic feature streaming wrapper (stream.py:33-254): same methods and result dictionary, the buffer is a _FrameRing."""

    def __init__(self, module, frame_buffer_size=1):
        self.module = module
        self.frame_buffer = None
        self.frame_buffer_size = frame_buffer_size
        self.start_time = None

    def reset_stream(self):
        self.stop_streaming()
        self.frame_buffer = _FrameRing()

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
        self.frame_buffer.push(frame, self.frame_buffer_size)
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
        feats = self.frame_buffer.stacked()
        times = np.array([self.get_elapsed_time()])
        # tools.dict_unsqueeze: a leading batch axis on every entry
        return {KEY_FEATS: feats.unsqueeze(0) if isinstance(feats, torch.Tensor) else feats[None], KEY_TIMES: times[None]}

    def get_elapsed_time(self, decimals=3):
        if self.start_time is None:
            return 0
        return round(_current_time(decimals) - self.start_time, decimals)


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
