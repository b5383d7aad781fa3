#!/usr/bin/env python
"""ncu target: the flash-attention kernel alone at the DiT's joint-attention shape (B=2, H=16, L=4442, D=64)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402

torch.manual_seed(0)
q = torch.randn(2, 4442, 16, 64, device="cuda").half()
k = torch.randn(2, 4442, 16, 64, device="cuda").half()
v = torch.randn(2, 4442, 16, 64, device="cuda").half()
for _ in range(4):
    ops.attention(q, k, v)
torch.cuda.synchronize()
print("prof attn done")
