import time, sys, numpy as np, torch
sys.path.insert(0,'.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
dev=torch.device('cuda',0)
for name,kw,sr,sec,B in (('HCQT',dict(sample_rate=22050,hop_length=256,n_bins=360,bins_per_octave=60),22050,30.0,16),('MelSpec',dict(),16000,20.0,64),('VQT',dict(sample_rate=22050,hop_length=512),22050,240.0,4)):
    m=getattr(ab,name)(device=dev,**kw)
    y=piano_like(int(sr*sec),sr,seed=1)
    a=torch.from_numpy(np.stack([y]*B)).to(dev)
    for _ in range(5): m.process_audio(a)
    torch.cuda.synchronize()
    for rep in range(3):
        t=time.perf_counter()
        for _ in range(50): o=m.process_audio(a)
        th=time.perf_counter()-t
        torch.cuda.synchronize()
        tt=time.perf_counter()-t
        print(name,'host enqueue per call %.3f ms, total per call %.3f ms'%(th/50*1e3, tt/50*1e3))
