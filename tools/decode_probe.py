#!/usr/bin/env python
"""Is the dense SDF decode host-bound or device-bound?  Times one 257^3 decode three ways: device time (CUDA events),
host time to ENQUEUE it (perf_counter before any synchronisation), and per-kernel-family device time of one chunk
(each family re-issued alone into a CUDA graph)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402
from r3g.pipelines import HUNYUAN3D_2_CONFIG  # noqa: E402
from r3g.vae import ShapeVAE  # noqa: E402

vae = ShapeVAE(device="cuda", **HUNYUAN3D_2_CONFIG["vae"]).init_random(1)
torch.manual_seed(0)
lat = (torch.randn(1, 3072, 1024, device="cuda") * 0.5).half()
R = 256
grid = torch.empty(1, R + 1, R + 1, R + 1, device="cuda", dtype=torch.float32)
bounds = [-1.01, -1.01, -1.01, 1.01, 1.01, 1.01]
geo = vae.geo_decoder
geo.decode_grid(lat, bounds, R, grid)
torch.cuda.synchronize()
out = {}
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    geo.decode_grid(lat, bounds, R, grid)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    out[f"rep{rep}"] = dict(device_ms=e0.elapsed_time(e1), host_enqueue_ms=1e3 * t_host)
    print(out[f"rep{rep}"], flush=True)

# one chunk, family by family
FAMILIES = ["linear", "attention", "layernorm", "lnpost_dot", "grid_fourier"]
calls = []
orig = {f: getattr(ops, f) for f in FAMILIES}


def wrap(name):
    def rec(*a, **kw):
        calls.append((name, a, kw))
        return orig[name](*a, **kw)
    return rec


for f in FAMILIES:
    setattr(ops, f, wrap(f))
try:
    ws = geo._workspace(65536)
    kv = geo._project_kv(lat)
    calls.clear()
    ops.grid_fourier(ws["emb"][:65536], 0, 65536, R, bounds, geo.num_freqs, geo.include_pi)
    geo._decode(ws, 65536, kv, grid.view(-1)[:65536])
finally:
    for f in FAMILIES:
        setattr(ops, f, orig[f])
torch.cuda.synchronize()


def time_graph(sel, reps=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for n, a, kw in sel:
            orig[n](*a, **kw)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for n, a, kw in sel:
            orig[n](*a, **kw)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out["chunk_graph_ms"] = time_graph(calls)
out["chunk_launches"] = len(calls)
for i, (n, a, kw) in enumerate(calls):
    shape = tuple(a[0].shape) if hasattr(a[0], "shape") else None
    out[f"k{i}_{n}"] = dict(ms=time_graph([(n, a, kw)]), x=shape, w=tuple(a[1].shape) if n == "linear" else None)
    print(f"k{i}_{n}", out[f"k{i}_{n}"], flush=True)
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("R3G_PROBE_OUT", "decode_probe.json")), "w"), indent=1)
