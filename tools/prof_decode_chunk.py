#!/usr/bin/env python
"""ncu target: the geo-decoder on grid chunks at the bench shape (3072 latents, 1024 wide): query_proj, ln_1, c_q,
cross attention (Lq = chunk, Lk = 3072), c_proj, ln_3, c_fc, mlp.c_proj, ln_post + output_proj."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g.pipelines import HUNYUAN3D_2_CONFIG  # noqa: E402
from r3g.vae import ShapeVAE  # noqa: E402

vae = ShapeVAE(**HUNYUAN3D_2_CONFIG["vae"]).init_random(seed=5)
torch.manual_seed(1)
lat = (torch.randn(1, 3072, 1024, device="cuda") * 0.5).half()
R = int(os.environ.get("R3G_PROF_OCTREE", "64"))      # 65^3 = 274 625 queries: a few chunks
grid = vae.volume_decoder(lat, vae.geo_decoder, bounds=1.01, octree_resolution=R)
grid = vae.volume_decoder(lat, vae.geo_decoder, bounds=1.01, octree_resolution=R)
torch.cuda.synchronize()
print("prof decode done", float(grid.abs().mean()))
