"""
Re-pins the golden fixtures on the REFERENCE ITSELF: runs the unmodified amt_tools.features modules (librosa-backed) of
/root/reference on the seeded synthetic audio of the five BASELINE.json configurations (at fixture sizes) and writes
tests/golden/golden_ref.npz with the versions of librosa / soxr / numpy / scipy inside.

    python tests/golden/regen_from_reference.py [--reference /root/reference] [--out tests/golden/golden_ref.npz]

It needs librosa (and, for CQT / VQT / HCQT with librosa >= 0.10, soxr).  Neither is installed in the build container nor in
its offline wheelhouse, so the committed goldens (golden_v1.npz, make_golden.py) are outputs of the oracle, and the CQT-family
stages of the oracle stay "parity unpinned" (oracle/librosa_stages.py).  The day librosa is importable this one command
replaces them: tests/test_oracle.py::test_reference_goldens_if_present and tests/test_gpu_parity.py::test_golden_fixtures
pick golden_ref.npz up automatically, check its provenance record, and hold BOTH the oracle and the CUDA path to it.

Only the `features` subpackage of the reference is imported (under a stub parent package): the rest of amt_tools pulls in
mir_eval, jams, sounddevice, ... which the feature path does not need.  The one symbol features/common.py takes from
amt_tools.tools is FLOAT32 (features/common.py:137).
"""
import argparse
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

# name: (reference class, kwargs, sample_rate, seconds, seed) -- the five BASELINE.json configurations at fixture sizes, dB and linear
CASES = {
    'c1_cqt192_db': ('CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24), 22050, 30.0, 0),
    'c1_cqt192_lin': ('CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24, decibels=False), 22050, 30.0, 0),
    'c2_mel_db': ('MelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048), 16000, 20.0, 1000),
    'c2_mel_lin': ('MelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048, decibels=False), 16000, 20.0, 1000),
    'c3_hcqt_db': ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 22050, 30.0, 2000),
    'c3_hcqt_lin': ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60, decibels=False), 22050, 30.0, 2000),
    'c4_stft_db': ('STFT', dict(sample_rate=22050, hop_length=512, n_fft=2048), 22050, 10.0, 3000),
    'c4_vqt_db': ('VQT', dict(sample_rate=22050, hop_length=512), 22050, 10.0, 3000),
    'c4_vqt_lin': ('VQT', dict(sample_rate=22050, hop_length=512, decibels=False), 22050, 10.0, 3000),
    'c4_power_db': ('SignalPower', dict(sample_rate=22050, hop_length=512), 22050, 10.0, 3000),
    'c5_hcqt_db_short': ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 22050, 6.0, 4000),
    'c5_mel_db_short': ('MelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048), 16000, 6.0, 4001),
}


def import_reference_features(reference_root):
    """amt_tools.features of the reference, without importing the rest of the package."""
    pkg_dir = os.path.join(reference_root, 'amt_tools')
    if not os.path.isdir(os.path.join(pkg_dir, 'features')):
        raise SystemExit('regen_from_reference: %s does not hold the reference (amt_tools/features missing)' % reference_root)
    parent = types.ModuleType('amt_tools')
    parent.__path__ = [pkg_dir]
    tools = types.ModuleType('amt_tools.tools')
    tools.FLOAT32 = 'float32'            # amt_tools/tools/constants.py
    parent.tools = tools
    sys.modules['amt_tools'] = parent
    sys.modules['amt_tools.tools'] = tools
    # features/__init__.py also imports the stream classes (sounddevice, ...): import the module files one by one instead
    feats = types.ModuleType('amt_tools.features')
    feats.__path__ = [os.path.join(pkg_dir, 'features')]
    sys.modules['amt_tools.features'] = feats
    out = {}
    for mod, names in (('common', ['FeatureModule']), ('waveform', ['WaveformWrapper']), ('stft', ['STFT']), ('mel', ['MelSpec']),
                       ('vqt', ['VQT']), ('cqt', ['CQT']), ('hvqt', ['HVQT']), ('hcqt', ['HCQT']), ('power', ['SignalPower']),
                       ('combo', ['FeatureCombo'])):
        m = importlib.import_module('amt_tools.features.' + mod)
        for n in names:
            out[n] = getattr(m, n)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_ref.npz'))
    args = ap.parse_args()
    try:
        import librosa
    except Exception as e:  # noqa: BLE001
        raise SystemExit('regen_from_reference: librosa is not importable here (%s).\nThe reference computes every feature through librosa '
                         '(features/stft.py:66, mel.py:64, vqt.py:183): without it there is no reference output to pin against, and the '
                         'committed goldens stay oracle outputs (tests/golden/make_golden.py).' % e)
    versions = {'librosa': librosa.__version__, 'numpy': np.__version__}
    for name in ('soxr', 'scipy', 'resampy'):
        try:
            versions[name] = importlib.import_module(name).__version__
        except Exception:  # noqa: BLE001
            versions[name] = 'absent'
    ref = import_reference_features(args.reference)
    from amt_tools_b200.synth import piano_like
    out = {}
    for case, (cls, kw, sr, sec, seed) in CASES.items():
        y = piano_like(int(sr * sec), sr, seed=seed)
        m = ref[cls](**kw)
        feats = np.asarray(m.process_audio(y))
        out[case] = np.ascontiguousarray(feats, dtype=np.float32)
        out[case + '__frames'] = np.int64(m.get_expected_frames(y))
        out[case + '__times'] = np.asarray(m.get_times(y), dtype=np.float64) if cls != 'VQT' else np.zeros(0)
        print('%-18s %-12s %s' % (case, cls, feats.shape))
    out['__provenance__'] = np.array(['reference'])
    out['__versions__'] = np.array(['%s=%s' % kv for kv in sorted(versions.items())])
    out['__cases__'] = np.array(['%s|%s|%r|%d|%r|%d' % (c, v[0], sorted(v[1].items()), v[2], v[3], v[4]) for c, v in CASES.items()])
    np.savez_compressed(args.out, **out)
    print('wrote %s  (%s)' % (args.out, ', '.join(out['__versions__'])))


if __name__ == '__main__':
    main()
