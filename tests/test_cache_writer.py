"""The cache writer's npz container (amt_tools_b200/precompute.py): written by hand for parallel deflate, it must stay what the
reference's loader reads -- `np.load` of a zip of .npy members with the keys fs, hop_length, features (datasets/common.py:245-250)."""
import zipfile

import numpy as np
import pytest

from amt_tools_b200 import precompute as pc


@pytest.mark.parametrize('shape', [(6, 36, 20001), (1, 229, 626), (3, 5, 0), (431,), (2, 3, 7)])
@pytest.mark.parametrize('compressed', [True, False])
def test_npz_round_trip(tmp_path, shape, compressed, monkeypatch):
    monkeypatch.setattr(pc, '_CHUNK', 1 << 20)        # several deflate pieces per array
    rng = np.random.RandomState(len(shape))
    feats = (rng.rand(*shape).astype(np.float32) ** 3) if int(np.prod(shape)) else np.zeros(shape, np.float32)
    path = str(tmp_path / 'Data' / 'feats' / 'track.npz')
    assert pc._write(path, 22050, 256, feats, compressed) == path
    with zipfile.ZipFile(path) as zf:
        assert zf.testzip() is None                   # CRCs and sizes of every member are right
        assert sorted(zf.namelist()) == ['features.npy', 'fs.npy', 'hop_length.npy']
    z = np.load(path)
    assert z[pc.KEY_FS].shape == () and z[pc.KEY_HOP].shape == ()        # 0-d, as np.savez stores them (the loader calls .item())
    assert z[pc.KEY_FS].item() == 22050 and z[pc.KEY_HOP].item() == 256
    assert z[pc.KEY_FEATS].dtype == np.float32 and np.array_equal(z[pc.KEY_FEATS], feats)
    assert not list(tmp_path.glob('**/*.tmp.*'))      # written next to the final name, then renamed


def test_deflated_file_matches_numpy_reader_on_fortran_and_int_inputs(tmp_path):
    a = np.asfortranarray(np.arange(24, dtype=np.int16).reshape(2, 3, 4))
    path = str(tmp_path / 'x.npz')
    with open(path, 'wb') as f:
        pc._savez_deflate(f, {'a': a, 'b': np.float64(3.5)}, 1)
    z = np.load(path)
    assert np.array_equal(z['a'], a) and z['b'].shape == () and z['b'].item() == 3.5
