"""
Streaming wrappers (amt_tools_b200/stream.py) against the reference's semantics
(/root/reference/amt_tools/features/stream.py:33-254, 637-779): buffer handling, slice boundaries, termination, and --
on the GPU -- that the batched look-ahead path returns exactly what one process_audio call per slice returns.
"""
import numpy as np
import pytest
import torch

import amt_tools_b200 as ab
from amt_tools_b200.stream import AudioStream, FeatureStream, KEY_FEATS, KEY_TIMES
from amt_tools_b200.synth import piano_like


class _FakeModule(object):
    """Host-only stand-in: one 'feature' per slice = (first sample, slice length); records what it was asked to process."""
    sample_rate, output = 100, 'numpy'

    def __init__(self, hop=4, need=7):
        self.hop, self.need, self.calls = hop, need, []

    def get_num_samples_required(self):
        return self.need

    def get_hop_length(self):
        return self.hop

    def get_num_channels(self):
        return 1

    def get_feature_size(self):
        return 2

    def _one(self, a):
        first = float(a[0]) if len(a) else -1.0
        return np.array([[[first], [float(len(a))]]], dtype=np.float32)

    def process_audio(self, audio):
        if isinstance(audio, list):
            self.calls.append(len(audio))
            return [self._one(a) for a in audio]
        self.calls.append(1)
        return self._one(audio)


def _reference_slices(n, hop, need):
    # stream.py:746-755 with query_finished of stream.py:777
    out, cs = [], 0
    while not cs > n:
        out.append((cs, min(n, cs + need) - cs))
        cs += hop
    return out


@pytest.mark.parametrize('lookahead', [1, 3, 64])
@pytest.mark.parametrize('n', [0, 1, 7, 8, 23, 24])
def test_audio_stream_slices_match_reference_loop(n, lookahead):
    m = _FakeModule()
    audio = np.arange(n, dtype=np.float32) + 1000
    st = AudioStream(m, audio=audio, lookahead=lookahead)
    assert st.extract_frame_features() is None   # not started yet (stream.py:734)
    st.start_streaming()
    got = []
    while not st.query_finished():
        f = st.extract_frame_features()
        got.append((int(f[0, 0, 0]) - 1000 if f[0, 1, 0] else len(audio), int(f[0, 1, 0])))
    assert got == _reference_slices(n, m.hop, m.need)
    assert st.extract_frame_features() is None
    assert max(m.calls, default=1) <= lookahead
    st.stop_streaming()
    assert not st.query_active()


def test_frame_buffer_semantics():
    m = _FakeModule()
    st = AudioStream(m, frame_buffer_size=3, audio=np.arange(40, dtype=np.float32))
    st.prime_frame_buffer(2)
    assert len(st.frame_buffer) == 2 and not st.query_frame_buffer_full()
    st.start_streaming()
    d = st.buffer_new_frame()
    assert d[KEY_FEATS].shape == (1, 1, 2, 3) and d[KEY_TIMES].shape == (1, 1)
    assert np.all(d[KEY_FEATS][0, :, :, :2] == 0)
    d = st.buffer_new_frame()   # buffer full: the oldest (empty) frame is dropped
    assert d[KEY_FEATS].shape == (1, 1, 2, 3) and len(st.frame_buffer) == 3
    assert d[KEY_FEATS][0, 0, 0, 1] == 0.0 and d[KEY_FEATS][0, 0, 0, 2] == 4.0
    st.reset_stream()
    assert st.frame_buffer == [] and st.current_sample == 0 and not st.query_active()
    base = FeatureStream(m)
    assert base.extract_frame_features() is NotImplementedError and base.query_finished() is NotImplementedError


def test_playback_is_rejected():
    with pytest.raises(ValueError):
        AudioStream(_FakeModule(), audio=np.zeros(4, dtype=np.float32), playback=True)


@pytest.mark.gpu
@pytest.mark.parametrize('name,kw', [
    ('MelSpec', dict(sample_rate=16000, hop_length=512)),
    ('CQT', dict(sample_rate=22050, hop_length=512, n_bins=84, bins_per_octave=12)),
    ('STFT', dict(sample_rate=16000, hop_length=512, decibels=False)),
])
def test_lookahead_equals_one_call_per_slice(name, kw):
    m = getattr(ab, name)(**kw)
    y = piano_like(int(kw['sample_rate'] * 1.3) + 17, kw['sample_rate'], seed=5)
    per, batched = AudioStream(m, audio=y), AudioStream(m, audio=y, lookahead=16)
    per.start_streaming()
    batched.start_streaming()
    n = 0
    while not per.query_finished():
        a, b = per.extract_frame_features(), batched.extract_frame_features()
        assert a.shape == b.shape and a.shape[:2] == (m.get_num_channels(), m.get_feature_size())
        assert torch.equal(a, b)   # same kernels, per-clip dB maximum: bit-identical
        n += 1
    assert batched.query_finished() and n == 1 + len(y) // m.get_hop_length()
    d = batched.buffer_empty_frame()
    assert d[KEY_FEATS].is_cuda and d[KEY_FEATS].shape[-1] == 1
