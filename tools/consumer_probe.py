"""Developer probe: does the upload of step i + 1 overlap the kernels of step i in the device-resident consumer loop?"""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
from amt_tools_b200.synth import piano_like
dev = torch.device('cuda', 0)
B = 8
mods = [ab.HCQT(22050, 256, n_bins=360, bins_per_octave=60, device=dev), ab.MelSpec(device=dev)]
host = [torch.from_numpy(np.stack([piano_like(sr * 240, sr, seed=i)] * B)).pin_memory() for i, sr in enumerate((22050, 16000))]
stage = [[torch.empty_like(h, device=dev) for h in host] for _ in range(2)]
up = torch.cuda.Stream(dev)
ms = [torch.cuda.Stream(dev) for _ in mods]
cur = torch.cuda.current_stream(dev)
for mode in ('upload_only', 'compute_only', 'both'):
    for m, d in zip(mods, stage[0]):
        m.process_audio(d)
    torch.cuda.synchronize()
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    t0 = time.perf_counter()
    host_t = []
    for i in range(12):
        slot = i % 2
        if mode != 'compute_only':
            with torch.cuda.stream(up):
                up.wait_event(consumed[slot])
                for d, h in zip(stage[slot], host):
                    d.copy_(h, non_blocking=True)
                uploaded[slot].record(up)
        if mode != 'upload_only':
            res = []
            for m, d, st in zip(mods, stage[slot], ms):
                st.wait_stream(cur)
                if mode == 'both':
                    st.wait_event(uploaded[slot])
                with torch.cuda.stream(st):
                    res.append(m.process_audio(d).reshape(B, -1).mean(dim=1))
            for st in ms:
                cur.wait_stream(st)
            consumed[slot].record(cur)
        host_t.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    total = time.perf_counter() - t0
    print(mode, 'total per step %.2f ms; host enqueue times (ms):' % (total / 12 * 1e3), ' '.join('%.1f' % (x * 1e3) for x in host_t))
