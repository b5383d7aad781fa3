#!/usr/bin/env python
"""Dense SDF decode (257^3 queries through the geo-decoder) timed against the chunk size: large chunks give the
GEMMs full waves, small chunks keep the [chunk, 1024] / [chunk, 4096] intermediates inside the 126 MB L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g.pipelines import HUNYUAN3D_2_CONFIG  # noqa: E402
from r3g.vae import ShapeVAE  # noqa: E402

vae = ShapeVAE(device="cuda", **HUNYUAN3D_2_CONFIG["vae"]).init_random(1)
torch.manual_seed(0)
lat = (torch.randn(1, 3072, 1024, device="cuda") * 0.5).half()
R = 256
grid = torch.empty(1, R + 1, R + 1, R + 1, device="cuda", dtype=torch.float32)
bounds = [-1.01, -1.01, -1.01, 1.01, 1.01, 1.01]
out = []
sizes = [int(x) for x in os.environ.get("R3G_CHUNKS", "8192,16384,24576,32768,49152,65536,131072").split(",")]
for cq in sizes:
    vae.geo_decoder.chunk_queries = cq
    vae.geo_decoder.decode_grid(lat, bounds, R, grid)  # warm-up (workspace allocation)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vae.geo_decoder.decode_grid(lat, bounds, R, grid)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = vae.geo_decoder.flops_per_query(3072) * (R + 1) ** 3
    out.append(dict(chunk=cq, ms=ms, tflops=fl / ms / 1e9))
    print(out[-1], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "decode_sweep.json"), "w"))
