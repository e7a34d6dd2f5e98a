"""
Generates tests/golden/golden_v1.npz: outputs of the float64 oracle (oracle/modules.py) on seeded synthetic
audio, for the module configurations of BASELINE.json at small sizes.  The reference itself (librosa-backed
amt_tools.features) cannot be imported in this container (no librosa / soxr), so these are oracle outputs,
not reference outputs -- see oracle/librosa_stages.py for the parity status.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import modules as om  # noqa: E402

CASES = {
    # name: (oracle ctor, kwargs, sample_rate, seconds, seed)
    'stft_db': ('OSTFT', dict(sample_rate=16000, hop_length=512, n_fft=2048), 16000, 2.0, 101),
    'stft_lin': ('OSTFT', dict(sample_rate=16000, hop_length=512, n_fft=2048, decibels=False), 16000, 2.0, 101),
    'mel_db': ('OMelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048), 16000, 2.0, 102),
    'mel_htk_lin': ('OMelSpec', dict(sample_rate=16000, hop_length=512, n_mels=229, n_fft=2048, htk=True, decibels=False), 16000, 2.0, 102),
    'cqt192_db': ('OCQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24), 22050, 2.0, 103),
    'cqt192_lin': ('OCQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24, decibels=False), 22050, 2.0, 103),
    'vqt84_lin': ('OVQT', dict(sample_rate=22050, hop_length=512, decibels=False), 22050, 2.0, 104),
    'hcqt_lin': ('OHCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60, decibels=False), 22050, 1.0, 105),
    'power_db': ('OSignalPower', dict(sample_rate=22050, hop_length=512), 22050, 2.0, 106),
}


def main():
    out = {}
    for name, (ctor, kw, sr, sec, seed) in CASES.items():
        y = piano_like(int(sr * sec), sr, seed=seed)
        feats = getattr(om, ctor)(dtype=np.float64, **kw).process_audio(y)
        out[name] = feats.astype(np.float32)
        print(name, feats.shape)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden_v1.npz'), **out)


if __name__ == '__main__':
    main()
