#!/usr/bin/env python
"""Short ncu target: the GEMM kernel at the DiT's shapes with their real epilogues.
Launch order (after 2 warm-up launches): [2] linear1 of a single block (GELU on the MLP columns + fused q/k RMS norm),
[3] linear2 (K = 5120, gate + residual), [4] img MLP2 of a double block (M 6144, N 1024, K 4096, gate + residual),
[5] txt QKV (M 2740, fused q/k norm), [6] linear1 with the plain bias epilogue (reference point)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402

torch.manual_seed(0)
H, Mh = 1024, 4096
g64 = (1 + 0.1 * torch.randn(64, device="cuda")).half()


def mk(M, N, K):
    return (torch.randn(M, K, device="cuda").half(), (torch.randn(N, K, device="cuda") * 0.02).half(),
            torch.zeros(N, device="cuda").half(), torch.empty(M, N, device="cuda", dtype=torch.float16))


x1, w1, b1, y1 = mk(8884, 7168, 1024)
x2, w2, b2, y2 = mk(8884, 1024, 5120)
x3, w3, b3, y3 = mk(6144, 1024, 4096)
x4, w4, b4, y4 = mk(2740, 3072, 1024)
gate = torch.randn(2, 1024, device="cuda").half()
qk1 = dict(mode=ops.QKN_RMS, q_col0=0, k_col0=H + Mh, cols=H, eps=1e-6, q_w=g64, k_w=g64)
qk4 = dict(mode=ops.QKN_RMS, q_col0=0, k_col0=H, cols=H, eps=1e-6, q_w=g64, k_w=g64)
for _ in range(2):
    ops.linear(x1, w1, b1, out=y1)
ops.linear(x1, w1, b1, out=y1, act=ops.ACT_GELU_TANH, act_cols=(H, H + Mh), qk_norm=qk1)
ops.linear(x2, w2, b2, out=y2, gate=gate, gate_rows=4442, residual=y2)
ops.linear(x3, w3, b3, out=y3, gate=gate, gate_rows=3072, residual=y3)
ops.linear(x4, w4, b4, out=y4, qk_norm=qk4)
ops.linear(x1, w1, b1, out=y1)
torch.cuda.synchronize()
print("done")
