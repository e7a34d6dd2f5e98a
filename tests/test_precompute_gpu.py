"""Bulk precompute + npz cache writer: reference cache format, sharded, bit-identical to process_audio."""
import os

import numpy as np
import pytest
import torch

import amt_tools_b200 as ab
from amt_tools_b200 import precompute
from amt_tools_b200.synth import piano_like

pytestmark = pytest.mark.gpu


def test_precompute_writes_reference_cache_format(tmp_path):
    tracks = {'track_%02d' % i: piano_like(16000 * (2 + i % 3), 16000, seed=80 + i) for i in range(7)}
    m = ab.MelSpec()
    got = {}
    for rank in range(2):     # two "ranks" on one GPU: their union must be the whole corpus
        got.update(precompute.precompute_features(tracks, m, str(tmp_path), 'Synth', rank=rank, world_size=2,
                                                  max_batch_seconds=6.0))
    assert sorted(got) == sorted(tracks)
    for name, path in got.items():
        assert path == os.path.join(str(tmp_path), 'Synth', 'MelSpec', name + '.npz') and os.path.exists(path)
        feats, fs, hop = precompute.load_features(path)
        assert fs == 16000 and hop == 512 and feats.dtype == np.float32
        want = m.process_audio(tracks[name]).cpu().numpy()
        assert feats.shape == want.shape and np.array_equal(feats, want)
    # second call: everything is a cache hit, nothing recomputed
    again = precompute.precompute_features(tracks, m, str(tmp_path), 'Synth')
    assert again == {}
    dev = precompute.precompute_features(tracks, m, str(tmp_path), 'Synth', overwrite=True, keep_on_device=True, compressed=False)
    name = sorted(tracks)[0]
    assert dev[name][1].is_cuda and torch.equal(dev[name][1].cpu(), torch.from_numpy(precompute.load_features(dev[name][0])[0]))


def _framify_reference(a, win, hop=1, pad=True):
    # restatement of amt_tools/tools/utils.py:2922-2984 (librosa.util.pad_center + chunking)
    T = a.shape[-1]
    pl = win // 2
    size = T + 2 * pl if pad else max(win, T)
    lpad = (size - T) // 2
    widths = [(0, 0)] * (a.ndim - 1) + [(lpad, size - T - lpad)]
    a = np.pad(a, widths)
    hops = (size - 2 * pl) // hop
    return np.concatenate([np.expand_dims(a[..., i * hop:i * hop + win], axis=-2) for i in range(hops)], axis=-2)


def test_device_framify_matches_reference_restated():
    rng = np.random.RandomState(0)
    for shape, win, hop, pad in [((2, 1, 192, 200), 9, 1, True), ((1, 192, 37), 9, 1, False), ((3, 50), 5, 2, True), ((4, 5), 9, 1, False)]:
        a = rng.randn(*shape).astype(np.float32)
        got = ab.framify_activations(torch.from_numpy(a).cuda(), win, hop, pad).cpu().numpy()
        want = _framify_reference(a, win, hop, pad)
        assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize('fused', ['1', '0'])
def test_pipeline_executor_matches_direct_calls(fused, monkeypatch):
    """amtfeat_pipeline_* (upload / compute / download streams over staging slots) returns bit-identical features, with the dB
    epilogue fused into the download (default: one kernel stores the finished features straight into the pinned host buffer) and
    with the in-place pass + copy (AMTFEAT_PIPE_FUSED=0); linear features (no epilogue) go through the copy engine either way."""
    monkeypatch.setenv('AMTFEAT_PIPE_FUSED', fused)
    import ctypes as C
    from amt_tools_b200 import _lib
    mods = [ab.MelSpec(16000), ab.HCQT(22050, 256, n_bins=120, bins_per_octave=24, harmonics=[0.5, 1, 2, 3]),
            ab.STFT(16000, decibels=False), ab.SignalPower(22050)]
    srs = [16000, 22050, 16000, 22050]
    B = 3
    jobs = []
    for rep in range(5):            # more submissions than slots: buffers are recycled
        for m, sr in zip(mods, srs):
            n = 4 * ((sr * 2 + 500 * rep) // 4)
            audio = np.stack([piano_like(n, sr, seed=900 + 10 * rep + b) for b in range(B)])
            jobs.append((m, n, audio))
    max_in = max(B * n for _, n, _ in jobs)
    max_out = max(B * int(np.prod(m._out_shape(n))) for m, n, _ in jobs)
    max_ws = max(int(_lib.lib.amtfeat_workspace_bytes(m._dev_plan.handle, B, _lib.i64_array([n] * B))) for m, n, _ in jobs)
    pipe = C.c_void_p()
    _lib.check(_lib.lib.amtfeat_pipeline_create(0, 2, max_in, max_out, max_ws, C.byref(pipe)))
    try:
        keep, tickets = [], []
        for m, n, audio in jobs:
            per = int(np.prod(m._out_shape(n)))
            h_in = torch.from_numpy(audio).pin_memory()
            h_out = torch.empty(B * per, dtype=torch.float32).pin_memory()
            t = C.c_int64(-1)
            _lib.check(_lib.lib.amtfeat_pipeline_submit(
                pipe, m._dev_plan.handle, h_in.data_ptr(), _lib.i64_array([b * n for b in range(B)]), _lib.i64_array([n] * B),
                _lib.i64_array([b * per for b in range(B)]), B, h_out.data_ptr(), h_in.numel(), h_out.numel(), C.byref(t)))
            keep.append((h_in, h_out))
            tickets.append(t.value)
        assert tickets == list(range(len(jobs)))
        _lib.check(_lib.lib.amtfeat_pipeline_wait(pipe, tickets[-1]))
        _lib.check(_lib.lib.amtfeat_pipeline_wait(pipe, -1))
        for (m, n, audio), (_, h_out) in zip(jobs, keep):
            want = m.process_audio(audio).cpu()
            assert torch.equal(h_out.view(want.shape), want)
        # oversized batches are refused, not truncated
        with pytest.raises(ValueError):
            _lib.check(_lib.lib.amtfeat_pipeline_submit(
                pipe, mods[0]._dev_plan.handle, keep[0][0].data_ptr(), _lib.i64_array([0]), _lib.i64_array([max_in + 4]),
                _lib.i64_array([0]), 1, keep[0][1].data_ptr(), max_in + 4, 4, None))
    finally:
        _lib.lib.amtfeat_pipeline_destroy(pipe)
