#!/usr/bin/env python
"""Kernel micro-benchmarks on one B200 (CUDA events, warm-up, L2 flush between timed launches)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402

FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    FLUSH.zero_()


def timeit(fn, iters=10, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = []
    torch.manual_seed(0)
    only = os.environ.get("R3G_MB_ONLY", "")
    for (M, N, K) in [] if only and "linear" not in only else [(8884, 7168, 1024), (8884, 1024, 5120), (6144, 3072, 1024), (6144, 4096, 1024),
                      (6144, 1024, 4096), (2740, 3072, 1024), (32768, 4096, 1024), (32768, 1024, 4096),
                      (32768, 1024, 1024)]:
        x = torch.randn(M, K, device="cuda").half()
        w = torch.randn(N, K, device="cuda").half()
        b = torch.randn(N, device="cuda").half()
        y = torch.empty(M, N, device="cuda", dtype=torch.float16)
        ms = timeit(lambda: ops.linear(x, w, b, out=y))
        ms_t = timeit(lambda: torch.nn.functional.linear(x, w, b))
        fl = 2.0 * M * N * K
        out.append(dict(op="linear", M=M, N=N, K=K, ms=ms, tflops=fl / ms / 1e9, torch_ms=ms_t,
                        torch_tflops=fl / ms_t / 1e9))
        print(out[-1], flush=True)
    for (B, H, Lq, Lk) in [] if only and "attention" not in only else [(2, 16, 4442, 4442), (1, 16, 3072, 3072), (1, 16, 65536, 3072)]:
        q = torch.randn(B, Lq, H, 64, device="cuda").half()
        k = torch.randn(B, Lk, H, 64, device="cuda").half()
        v = torch.randn(B, Lk, H, 64, device="cuda").half()
        o = torch.empty_like(q)
        ms = timeit(lambda: ops.attention(q, k, v, out=o))
        qt, kt, vt = (t.transpose(1, 2).contiguous() for t in (q, k, v))
        ms_t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qt, kt, vt))
        fl = 4.0 * B * H * Lq * Lk * 64
        out.append(dict(op="attention", B=B, H=H, Lq=Lq, Lk=Lk, ms=ms, tflops=fl / ms / 1e9, torch_ms=ms_t,
                        torch_tflops=fl / ms_t / 1e9))
        print(out[-1], flush=True)
    for n in [] if only and "mc" not in only else (257, 513):
        ax = torch.linspace(-1.01, 1.01, n, device="cuda")
        x, y, z = torch.meshgrid(ax, ax, ax, indexing="ij")
        vol = (0.25 - torch.sqrt((torch.sqrt(x * x + y * y) - 0.6) ** 2 + z * z)).contiguous()
        del x, y, z
        ms = timeit(lambda: ops.marching_cubes(vol, 0.0), iters=5)
        v, f = ops.marching_cubes(vol, 0.0)
        out.append(dict(op="marching_cubes", n=n, ms=ms, verts=len(v), faces=len(f),
                        grid_gbs=n ** 3 * 4 * 3 / ms / 1e6))
        print(out[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("R3G_MB_OUT", "microbench.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
