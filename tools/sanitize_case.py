"""Small ragged batches through every kernel family, for compute-sanitizer (memcheck / racecheck) on the GPU box."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like

y = piano_like(22050 * 3 + 77, 22050, seed=1)
mods = (ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60), ab.CQT(22050, 128, n_bins=96, bins_per_octave=12), ab.MelSpec(), ab.VQT(22050, 512),
        ab.HCQT(22050, 512, harmonics=[0.25, 0.5, 1], n_bins=96, bins_per_octave=24))
for m in mods:
    out = m.process_audio([y, y[:20011], y[:700], y[:23807]])
    torch.cuda.synchronize()
    print(type(m).__name__, [tuple(o.shape) for o in out], float(out[0].max()))

# round 2: device ingest (PCM conversion, resampler, rms_norm), the pipelined executor with the fused download epilogue,
# framify, STFT / SignalPower / WaveformWrapper
from amt_tools_b200 import ingest, precompute  # noqa: E402
dev = torch.device('cuda', 0)
pcm = torch.from_numpy((y[:30001] * 8000).astype(np.int16)).to(dev)
f = ingest.pcm16_to_float(pcm, 1.0 / 32768.0, device=dev)
r = ingest.resample([f, f[:4001]], 22050, 16000, device=dev)
z = ingest.rms_norm(r[0], device=dev)
torch.cuda.synchronize()
print('ingest', tuple(f.shape), [tuple(q.shape) for q in r], float(z.abs().max()))
for m in (ab.STFT(16000, 512), ab.SignalPower(22050), ab.WaveformWrapper(22050, 512, win_length=1024)):
    out = m.process_audio([y, y[:5003]])
    torch.cuda.synchronize()
    print(type(m).__name__, [tuple(o.shape) for o in out])
fr = ab.framify_activations(torch.rand(3, 50, device=dev), 9)
torch.cuda.synchronize()
print('framify', tuple(fr.shape))
import tempfile  # noqa: E402
with tempfile.TemporaryDirectory() as d:
    tracks = {'t%d' % i: y[:22050 + 1000 * i] for i in range(5)}
    for m in (ab.MelSpec(), ab.HCQT(22050, 512, harmonics=[0.5, 1], n_bins=48, bins_per_octave=12)):
        res = precompute.precompute_features(tracks, m, d, 'San', max_batch_seconds=2.5, writers=2)
        print('precompute', type(m).__name__, len(res))
