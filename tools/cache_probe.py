"""Developer probe: the precompute + cache leg of bench.py alone, with its stage timings."""
import sys, json
import torch
sys.path.insert(0, '.')
import amt_tools_b200 as ab
import bench
desc, spec, seconds, B = bench.WORKLOADS['c5']
dev = torch.device('cuda', 0)
for compressed, tracks in ((False, 16), (True, 4), (True, 16)):
    r = bench.cache_leg(ab, torch, dev, spec, seconds, tracks, compressed, 15)
    print(json.dumps({k: v for k, v in r.items() if k != 'dir'}))
