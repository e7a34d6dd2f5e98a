"""Developer script (GPU box): one very long clip (large sample / frame indices) -- interior frames must equal those computed from a
short excerpt around them (linear features; the excerpt is long enough for every filter and decimator to settle)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like

sr, hours = 22050, 2.0
n = int(sr * 3600 * hours)
rng = np.random.RandomState(1)
seg = piano_like(sr * 60, sr, seed=2)
y = np.tile(seg, n // len(seg) + 1)[:n].copy()
y *= (1.0 + 0.1 * np.sin(np.arange(n, dtype=np.float32) * 1e-6)).astype(np.float32)     # not periodic
dev = torch.device('cuda', 0)
yd = torch.from_numpy(y).to(dev)
bad = 0
for name, kw in (('STFT', dict(sample_rate=sr, hop_length=512, n_fft=2048)), ('MelSpec', dict(sample_rate=sr, hop_length=512)),
                 ('VQT', dict(sample_rate=sr, hop_length=512)), ('HCQT', dict(sample_rate=sr, hop_length=256, n_bins=360, bins_per_octave=60)),
                 ('SignalPower', dict(sample_rate=sr, hop_length=512))):
    m = getattr(ab, name)(decibels=False, **kw)
    full = m.process_audio(yd)
    T = full.shape[-1]
    hop = kw['hop_length']
    assert T == m.get_expected_frames(y), (T, m.get_expected_frames(y))
    worst = 0.0
    for frac in (0.37, 0.93, 0.999):
        t = int(T * frac) // 64 * 64            # frame index; its sample position t * hop is a multiple of every level's hop
        half = 12 * sr // hop * hop             # 12 s either side
        a, b = t * hop - half, t * hop + half
        if b > n:
            b = n
        sub = m.process_audio(yd[a:b].contiguous())
        k = half // hop
        fa = full[..., t - 20:t + 20].double().cpu().numpy()
        fs = sub[..., k - 20:k + 20].double().cpu().numpy()
        worst = max(worst, np.abs(fa - fs).max() / max(np.abs(fa).max(), 1e-30))
    ok = worst < 2e-5 and bool(torch.isfinite(full).all())
    bad += not ok
    print('%s %-12s n=%d T=%d out=%.2f GB  interior frames vs excerpt: max|d|/peak %.2e' % ('ok ' if ok else 'BAD', name, n, T, full.numel() * 4 / 1e9, worst), flush=True)
    del full
    torch.cuda.empty_cache()
sys.exit(1 if bad else 0)
