#!/usr/bin/env python
"""Short target for `ncu --set full`: the hot kernels at the metric's real shapes, a few launches each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402

torch.manual_seed(0)
x = torch.randn(8884, 1024, device="cuda").half()
w1 = (torch.randn(7168, 1024, device="cuda") * 0.02).half()
b1 = torch.zeros(7168, device="cuda").half()
y = torch.empty(8884, 7168, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.linear(x, w1, b1, out=y, act=ops.ACT_GELU_TANH, act_cols=(1024, 5120))
q = torch.randn(2, 4442, 16, 64, device="cuda").half()
k = torch.randn(2, 4442, 16, 64, device="cuda").half()
v = torch.randn(2, 4442, 16, 64, device="cuda").half()
for _ in range(3):
    ops.attention(q, k, v)
n = 513
ax = torch.linspace(-1.01, 1.01, n, device="cuda")
gx, gy, gz = torch.meshgrid(ax, ax, ax, indexing="ij")
vol = (0.25 - torch.sqrt((torch.sqrt(gx * gx + gy * gy) - 0.6) ** 2 + gz * gz)).contiguous()
del gx, gy, gz
for _ in range(2):
    ops.marching_cubes(vol, 0.0)
d = torch.rand(64, 1022, 1022, device="cuda") + 0.5
import numpy as np
E = np.tile(np.eye(3, 4, dtype=np.float32), (64, 1, 1))
K = np.tile(np.array([[800, 0, 511], [0, 800, 511], [0, 0, 1]], np.float32), (64, 1, 1))
for _ in range(2):
    ops.unproject(d, E, K, torch.float64)
torch.cuda.synchronize()
print("prof target done")
