#!/usr/bin/env python
"""ncu target: ONE eager DiT forward at the bench shape (B = 2 CFG batch, 1370 + 3072 tokens, 16 + 32 blocks) after a
warm-up forward.  Used for
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:linear_kernel
      -> per-launch DRAM traffic of the GEMM family (bench.py roofline.traffic), and
  ncu --set full -k regex:<kernel> -s <skip> -c 1 -> the full section set of one launch of a shipped kernel."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g.dit import Hunyuan3DDiT  # noqa: E402
from r3g.pipelines import HUNYUAN3D_2_CONFIG  # noqa: E402

model = Hunyuan3DDiT(device="cuda", **HUNYUAN3D_2_CONFIG["model"]).init_random(0)
torch.manual_seed(0)
x = torch.randn(2, 3072, 64, device="cuda").half()
t = torch.full((2,), 0.5, device="cuda", dtype=torch.float16)
cond = {"main": torch.randn(2, 1370, 1536, device="cuda").half()}
n = int(os.environ.get("R3G_PROF_FORWARDS", "2"))
for _ in range(n):
    model(x, t, cond)
torch.cuda.synchronize()
print("prof dit forward done")
