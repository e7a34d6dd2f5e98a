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
