"""compute-sanitizer case for the chunked long-track path (range_max_kernel / range_finish_kernel / the dB epilogue with two loads in flight):
compute-sanitizer --tool memcheck python tools/sanitize_longtrack.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200 import longtrack as lt
from amt_tools_b200.synth import piano_like

dev = torch.device('cuda', 0)
for name, kw, sec, chunk in (('MelSpec', dict(sample_rate=16000, hop_length=512, n_fft=2048), 12, 128),
                             ('SignalPower', dict(sample_rate=16000, hop_length=512), 12, 128),
                             ('VQT', dict(sample_rate=22050, hop_length=512), 40, 640)):
    m = getattr(ab, name)(device=dev, **kw)
    y = torch.from_numpy(piano_like(kw['sample_rate'] * sec + 3, kw['sample_rate'], seed=3)).to(dev)
    whole = m.process_audio(y)
    got = lt.process_long_audio(m, y, chunk_frames=chunk)
    torch.cuda.synchronize()
    print(name, tuple(got.shape), float((got - whole).abs().max()))
