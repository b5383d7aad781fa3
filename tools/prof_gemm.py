#!/usr/bin/env python
"""Short ncu target: the GEMM kernel at one DiT shape, plain bias epilogue and GELU epilogue."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops
torch.manual_seed(0)
x = torch.randn(8884, 1024, device="cuda").half()
w = (torch.randn(7168, 1024, device="cuda") * 0.02).half()
b = torch.zeros(7168, device="cuda").half()
y = torch.empty(8884, 7168, device="cuda", dtype=torch.float16)
for _ in range(3):
    ops.linear(x, w, b, out=y)
for _ in range(2):
    ops.linear(x, w, b, out=y, act=ops.ACT_GELU_TANH, act_cols=(1024, 5120))
torch.cuda.synchronize()
print("done")
