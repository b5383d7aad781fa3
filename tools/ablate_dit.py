#!/usr/bin/env python
"""Where one DiT forward (B=2 CFG batch) spends its time, measured in-graph: the r3g calls of one eager forward are
recorded and re-issued into CUDA graphs -- all of them, all but one kernel family, and one family alone -- and each
graph is timed with CUDA events.  `minus` (full - without the family) is the family's marginal cost in its real
neighbourhood (cache state, launch gaps); `alone` is its cost back to back with itself."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))
from r3g import ops  # noqa: E402
from r3g.dit import Hunyuan3DDiT  # noqa: E402
from r3g.pipelines import HUNYUAN3D_2_CONFIG  # noqa: E402

FAMILIES = ["linear", "linear_pair", "attention", "layernorm", "qk_norm_", "gemv", "timestep_embedding"]


def time_graph(calls, reps=5):
    if not calls:
        return 0.0
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for fn, a, kw in calls:
            fn(*a, **kw)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for fn, a, kw in calls:
            fn(*a, **kw)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    model = Hunyuan3DDiT(device="cuda", **HUNYUAN3D_2_CONFIG["model"]).init_random(0)
    torch.manual_seed(0)
    x = torch.randn(2, 3072, 64, device="cuda").half()
    t = torch.full((2,), 0.5, device="cuda", dtype=torch.float16)
    cond = {"main": torch.randn(2, 1370, 1536, device="cuda").half()}
    model.group_streams = os.environ.get("R3G_GROUP", "1") != "0"
    model(x, t, cond)
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
        model(x, t, cond)
    finally:
        for f in FAMILIES:
            setattr(ops, f, orig[f])
    torch.cuda.synchronize()
    mk = lambda pred: [(orig[n], a, kw) for n, a, kw in calls if pred(n)]  # noqa: E731
    full = time_graph(mk(lambda n: True))
    out = {"full_ms": full, "launches": len(calls), "families": {}}
    for f in FAMILIES:
        n = sum(1 for c in calls if c[0] == f)
        if not n:
            continue
        minus = full - time_graph(mk(lambda m: m != f))
        alone = time_graph(mk(lambda m: m == f))
        out["families"][f] = {"launches": n, "minus_ms": minus, "alone_ms": alone}
        print(f, out["families"][f], flush=True)
    print(json.dumps(out))
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("R3G_ABLATE_OUT", "ablate_dit.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
