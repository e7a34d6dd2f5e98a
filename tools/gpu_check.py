"""Developer script: parity numbers of every module against the float64 oracle (run on the GPU box)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import modules as om  # noqa: E402


def compare(name, got, want, decibels, power=False):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    res = {'name': name, 'shape': list(got.shape), 'shape_ok': got.shape == want.shape}
    if got.shape != want.shape:
        res['want_shape'] = list(want.shape)
        return res
    if decibels:
        scale = 1.0 if power else 80.0
        d = np.abs(got - want) * scale
        top = want > ((-60.0) if power else 0.25)
        res['db_maxabs'] = float(d.max())
        res['db_maxabs_above_-60dB'] = float(d[top].max()) if top.any() else 0.0
    else:
        res['rel_l2'] = float(np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-30))
        res['max_rel_to_peak'] = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
    return res


def main():
    out = []
    y22 = piano_like(22050 * 10, 22050, seed=3)
    y16 = piano_like(16000 * 10, 16000, seed=4)
    cases = []
    for db in (False, True):
        cases += [
            ('STFT', ab.STFT(decibels=db), om.OSTFT(decibels=db), y16),
            ('STFT512', ab.STFT(decibels=db, n_fft=512, hop_length=128), om.OSTFT(decibels=db, n_fft=512, hop_length=128), y16),
            ('STFT-nc', ab.STFT(decibels=db, center=False), om.OSTFT(decibels=db, center=False), y16[:50001]),
            ('Mel', ab.MelSpec(decibels=db), om.OMelSpec(decibels=db), y16),
            ('Mel-htk', ab.MelSpec(decibels=db, htk=True), om.OMelSpec(decibels=db, htk=True), y16),
            ('Power', ab.SignalPower(22050, decibels=db), om.OSignalPower(22050, decibels=db), y22),
            ('CQT192', ab.CQT(22050, 512, db, n_bins=192, bins_per_octave=24), om.OCQT(22050, 512, db, n_bins=192, bins_per_octave=24), y22),
            ('VQT84', ab.VQT(22050, 512, db), om.OVQT(22050, 512, db), y22),
            ('HCQT', ab.HCQT(22050, 256, db, n_bins=360, bins_per_octave=60), om.OHCQT(22050, 256, db, n_bins=360, bins_per_octave=60), y22),
        ]
    cases.append(('Frames', ab.WaveformWrapper(22050, 512, win_length=1024), om.OWaveformWrapper(22050, 512, win_length=1024), y22[:30000]))
    for name, m, o, y in cases:
        t = time.time()
        got = m.process_audio(y)
        torch.cuda.synchronize()
        dt = time.time() - t
        want = o.process_audio(y)
        r = compare(name + ('-dB' if m.decibels else ''), got.cpu().numpy(), want, m.decibels, power=isinstance(m, ab.SignalPower))
        r['first_call_s'] = round(dt, 4)
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'gpu_check.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
