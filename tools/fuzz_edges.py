"""Developer script (GPU box): edge lengths (1, 2, hop - 1, hop, hop + 1, n_fft - 1, n_fft, n_fft + 1, tile boundaries, silence, DC, impulse)
in one ragged batch per module, CUDA path against the float64 oracle.  Exits non-zero on a miss."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amt_tools_b200 as ab  # noqa: E402
from amt_tools_b200.synth import piano_like  # noqa: E402
from oracle import modules as om  # noqa: E402

CASES = [
    ('STFT', dict(sample_rate=16000, hop_length=512, n_fft=2048)),
    ('STFT', dict(sample_rate=16000, hop_length=100, n_fft=256, win_length=200)),
    ('STFT', dict(sample_rate=16000, hop_length=512, n_fft=2048, center=False)),
    ('MelSpec', dict(sample_rate=16000, hop_length=512, n_fft=2048, n_mels=229)),
    ('MelSpec', dict(sample_rate=22050, hop_length=256, n_fft=1024, n_mels=80, center=False)),
    ('SignalPower', dict(sample_rate=22050, hop_length=512)),
    ('SignalPower', dict(sample_rate=22050, hop_length=256, win_length=1024, center=False)),
    ('WaveformWrapper', dict(sample_rate=22050, hop_length=512, win_length=1024)),
    ('CQT', dict(sample_rate=22050, hop_length=512, n_bins=192, bins_per_octave=24)),
    ('VQT', dict(sample_rate=22050, hop_length=512)),
    ('HCQT', dict(sample_rate=22050, hop_length=256, n_bins=360, bins_per_octave=60)),
    ('HVQT', dict(sample_rate=22050, hop_length=512, harmonics=[0.5, 1, 2], n_bins=72, bins_per_octave=12)),
]


def main():
    bad = 0
    rng = np.random.RandomState(0)
    for name, kw in CASES:
        hop = kw['hop_length']
        nfft = kw.get('n_fft') or kw.get('win_length') or 2048
        sr = kw['sample_rate']
        lens = sorted({1, 2, 3, hop - 1, hop, hop + 1, nfft - 1, nfft, nfft + 1, 8 * hop - 1, 8 * hop, 32 * hop + 1, 127 * hop + 5, int(sr * 1.37)})
        if not kw.get('center', True):
            lens = [n for n in lens if n >= (kw.get('win_length') or nfft)]
        base = piano_like(max(lens), sr, seed=3)
        clips = [base[:n].copy() for n in lens]
        clips.append(np.zeros(4 * hop + 7, dtype=np.float32))                       # silence
        clips.append(np.full(4 * hop + 7, 0.25, dtype=np.float32))                  # DC
        imp = np.zeros(6 * hop + 3, dtype=np.float32); imp[3 * hop + 1] = 1.0       # impulse
        clips.append(imp)
        clips.append((rng.randn(5 * hop + 11) * 1e-4).astype(np.float32))           # faint noise
        for decibels in ((False,) if name == 'WaveformWrapper' else (False, True)):
            mk = lambda mod, pre: getattr(mod, pre + name)(decibels=decibels, **{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()})
            m, o = mk(ab, ''), mk(om, 'O')
            worst = 0.0
            note = ''
            # clips the reference rejects (librosa: "Input signal length=... is too short for ...-octave CQT"): the module must raise too
            use = []
            for c in clips:
                try:
                    o.process_audio(c)
                    use.append(c)
                except ValueError:
                    try:
                        m.process_audio(c)
                        note += ' NO ERROR at n=%d (the oracle raises);' % len(c)
                    except ValueError:
                        pass
            clips_ok = use
            got = m.process_audio(clips_ok)
            for g, c in zip(got, clips_ok):
                g = g.cpu().numpy().astype(np.float64)
                w = np.asarray(o.process_audio(c), np.float64)
                if g.shape != w.shape:
                    if len(c) == 0 or w.size == 0 or g.size == 0:
                        continue
                    note += ' SHAPE %s vs %s at n=%d;' % (g.shape, w.shape, len(c))
                    continue
                if not w.size:
                    continue
                if not np.isfinite(g).all():
                    note += ' NONFINITE at n=%d;' % len(c)
                    continue
                if decibels:
                    scale, thr = (1.0, -60.0) if name == 'SignalPower' else (80.0, 0.25)
                    d = np.abs(g - w) * scale
                    top = w > thr
                    e = d[top].max() if top.any() else 0.0
                    lim = 1e-3 if name not in ('CQT', 'VQT', 'HCQT', 'HVQT') else 3e-3     # (one-frame clips: float32 itself is 1e-3 .. 3e-3 off)
                    if e > lim or d.max() > 3e-2:
                        note += ' dB top %.2e all %.2e at n=%d;' % (e, d.max(), len(c))
                    worst = max(worst, e)
                else:
                    den = max(np.linalg.norm(w), 1e-30)
                    e = np.linalg.norm(g - w) / den if np.abs(w).max() > 1e-12 else np.abs(g).max()
                    if e > 1e-5:
                        note += ' lin %.2e at n=%d;' % (e, len(c))
                    worst = max(worst, e)
            bad += bool(note)
            print('%s %-15s dB=%d %s worst=%.2e%s' % ('BAD' if note else 'ok ', name, decibels, kw, worst, note), flush=True)
    print('misses:', bad)
    return bad


if __name__ == '__main__':
    sys.exit(1 if main() else 0)
