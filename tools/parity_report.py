"""Developer script (GPU box): where the CQT-family dB error sits -- per octave, interior / tail, float64 and float32 oracle."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import modules as om  # noqa: E402


def breakdown(name, got, want, bpo, hop, sr, tail_s=4.5):
    """got / want: (C, F, T) dB features scaled to [0, 1]."""
    res = {'name': name}
    d = np.abs(got.astype(np.float64) - want) * 80.0
    top = want > 0.25
    res['all'] = float(d.max())
    res['top'] = float(d[top].max()) if top.any() else 0.0
    tail = int(np.ceil(tail_s * sr / hop))
    dt, tt = d[..., :-tail], top[..., :-tail]
    res['top_interior'] = float(dt[tt].max()) if tt.any() else 0.0
    per = []
    for c in range(got.shape[0]):
        row = []
        for o in range(0, got.shape[1], bpo):
            dd, tp = d[c, o:o + bpo, :-tail], top[c, o:o + bpo, :-tail]
            row.append(round(float(dd[tp].max()) if tp.any() else 0.0, 6))
        per.append(row)
    res['top_interior_per_channel_octave(low->high)'] = per
    return res


def main():
    sec = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
    y = piano_like(int(22050 * sec), 22050, seed=3)
    cases = [
        ('CQT192', 'CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24), 24),
        ('VQT84', 'VQT', dict(sample_rate=22050, hop_length=512), 12),
        ('HCQT', 'HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60), 60),
    ]
    out = []
    for label, ctor, kw, bpo in cases:
        got = getattr(ab, ctor)(**kw).process_audio(y).cpu().numpy()
        w64 = getattr(om, 'O' + ctor)(**kw).process_audio(y)
        w32 = getattr(om, 'O' + ctor)(dtype=np.float32, **kw).process_audio(y)
        for tag, want in (('gpu_vs_f64', w64), ('gpu_vs_f32', w32)):
            r = breakdown(label + ' ' + tag, got, np.asarray(want, np.float64), bpo, kw['hop_length'], 22050)
            print(json.dumps(r), flush=True)
            out.append(r)
        r = breakdown(label + ' f32_vs_f64', np.asarray(w32, np.float64), np.asarray(w64, np.float64), bpo, kw['hop_length'], 22050)
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'parity_report.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
