#!/usr/bin/env python
"""Where do 118 ms of 'marching cubes' go on a noise-like 257^3 grid?  GPU events vs host timers."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops
torch.manual_seed(0)
n = 257
c = torch.randn(1, 1, 40, 40, 40, device="cuda")
vol = torch.nn.functional.interpolate(c, size=(n, n, n), mode="trilinear", align_corners=True)[0, 0].contiguous()
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    v, f = ops.marching_cubes(vol, 0.0, bounds=[-1.01] * 3 + [1.01] * 3)
    b.record()
    torch.cuda.synchronize()
    print(f"iter {it}: gpu {a.elapsed_time(b):.2f} ms, host {1e3 * (time.perf_counter() - t0):.2f} ms, V={len(v)} F={len(f)}", flush=True)
