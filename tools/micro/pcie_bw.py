"""Pinned host <-> device copy bandwidth on this box (context for the e2e number in bench.py)."""
import torch
n = 742 * 1000 * 1000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device='cuda')
h2 = torch.empty(146 * 1000 * 1000 // 4, dtype=torch.float32).pin_memory()
d2 = torch.empty_like(h2, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: h.copy_(d, non_blocking=True)); print('D2H 742 MB: %.2f ms  %.1f GB/s' % (ms, 0.742 / ms * 1e3))
ms = t(lambda: d.copy_(h, non_blocking=True)); print('H2D 742 MB: %.2f ms  %.1f GB/s' % (ms, 0.742 / ms * 1e3))
def both():
    cur = torch.cuda.current_stream()
    s1.wait_stream(cur); s2.wait_stream(cur)
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    cur.wait_stream(s1); cur.wait_stream(s2)
ms = t(both); print('D2H 742 MB || H2D 146 MB: %.2f ms  D2H %.1f GB/s' % (ms, 0.742 / ms * 1e3))
