"""ORACLE -- test infrastructure only.  The reference's CPU path timed on the host cores (bench.py's
`cpu_baseline` object and its `--impl reference` arm).

The reference modules themselves cannot travel to the GPU box (/root/reference is absent there, and it needs
trimesh / skimage / diffusers), so `kind` is "port": the fp32 restatements of oracle/hy3d_ref.py (pinned to the
reference's own modules by tests/golden) and the C marching-cubes oracle.  A whole object costs about an hour
of CPU time at 256^3, so a BOUNDED SAMPLE of the same workload is timed and scaled
linearly; the sample and the scale factors are reported.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import hy3d_ref as R  # noqa: E402
import mc as omc  # noqa: E402


def _rand_sd_dit(depth, depth_single, H=1024, Mh=4096, ctx=1536, seed=0):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(n, o, i):
        sd[n + ".weight"] = torch.randn(o, i, generator=g) * 0.02
        sd[n + ".bias"] = torch.randn(o, generator=g) * 0.01

    lin("latent_in", H, 64); lin("time_in.in_layer", H, 256); lin("time_in.out_layer", H, H); lin("cond_in", H, ctx)
    for i in range(depth):
        p = f"double_blocks.{i}."
        for s in ("img", "txt"):
            lin(p + f"{s}_mod.lin", 6 * H, H); lin(p + f"{s}_attn.qkv", 3 * H, H); lin(p + f"{s}_attn.proj", H, H)
            lin(p + f"{s}_mlp.0", Mh, H); lin(p + f"{s}_mlp.2", H, Mh)
            sd[p + f"{s}_attn.norm.query_norm.scale"] = torch.ones(64)
            sd[p + f"{s}_attn.norm.key_norm.scale"] = torch.ones(64)
    for i in range(depth_single):
        p = f"single_blocks.{i}."
        lin(p + "modulation.lin", 3 * H, H); lin(p + "linear1", 3 * H + Mh, H); lin(p + "linear2", H, H + Mh)
        sd[p + "norm.query_norm.scale"] = torch.ones(64)
        sd[p + "norm.key_norm.scale"] = torch.ones(64)
    lin("final_layer.adaLN_modulation.1", 2 * H, H); lin("final_layer.linear", 64, H)
    return sd


def _rand_sd_geo(W=1024, seed=1):
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def lin(n, o, i, bias=True):
        sd[n + ".weight"] = torch.randn(o, i, generator=g) * 0.02
        if bias:
            sd[n + ".bias"] = torch.randn(o, generator=g) * 0.01

    def ln(n, d):
        sd[n + ".weight"], sd[n + ".bias"] = torch.ones(d), torch.zeros(d)

    p = "geo_decoder."
    lin(p + "query_proj", W, 51)
    ca = p + "cross_attn_decoder."
    ln(ca + "ln_1", W); ln(ca + "ln_2", W); ln(ca + "ln_3", W)
    lin(ca + "attn.c_q", W, W, False); lin(ca + "attn.c_kv", 2 * W, W, False)
    ln(ca + "attn.attention.q_norm", 64); ln(ca + "attn.attention.k_norm", 64)
    lin(ca + "attn.c_proj", W, W); lin(ca + "mlp.c_fc", 4 * W, W); lin(ca + "mlp.c_proj", W, 4 * W)
    ln(p + "ln_post", W); lin(p + "output_proj", 1, W)
    return sd


def time_object_sample(octree_resolution=256, steps=50, mc_grid=129, dit_reps=3, chunk_reps=4):
    """Times the sample and returns (objects_per_second, details).  Sample (all fp32, true widths):
      DiT: `dit_reps` x (DoubleStreamBlock + SingleStreamBlock) forward at B=2, L=1370+3072, averaged
                                                                                       -> x (16, 32) blocks x `steps`
      decode: `chunk_reps` chunks of 16000 grid queries through the geo-decoder, averaged -> x ceil((R+1)^3 / 16000)
      marching cubes: the C oracle on a `mc_grid`^3 sphere                            -> x ((R+1)/mc_grid)^3
    The ShapeVAE transformer (3.5 s of ~1 h) and the conditioner are left out of the sample."""
    torch.manual_seed(0)
    sd = _rand_sd_dit(1, 1)
    x = torch.randn(2, 3072, 64)
    c = torch.randn(2, 1370, 1536)
    img, txt = R._lin(sd, "latent_in", x), R._lin(sd, "cond_in", c)
    vec = torch.randn(2, 1024) * 0.1
    with torch.no_grad():
        t_double = t_single = 0.0
        for _ in range(dit_reps):
            t0 = time.perf_counter()
            img2, txt2 = R.double_block(sd, "double_blocks.0.", img, txt, vec, 16)
            t_double += (time.perf_counter() - t0) / dit_reps
            t0 = time.perf_counter()
            R.single_block(sd, "single_blocks.0.", torch.cat((txt2, img2), 1), vec, 16)
            t_single += (time.perf_counter() - t0) / dit_reps
        geo = _rand_sd_geo()
        lat = torch.randn(1, 3072, 1024) * 0.5
        q = torch.rand(1, 16000, 3) * 2 - 1
        t_chunk = 0.0
        for _ in range(chunk_reps):
            t0 = time.perf_counter()
            R.geo_decoder(geo, q, lat, 16, 8, False)
            t_chunk += (time.perf_counter() - t0) / chunk_reps
    n = mc_grid
    ax = np.linspace(-1.01, 1.01, n, dtype=np.float32)
    xx, yy, zz = np.meshgrid(ax, ax, ax, indexing="ij")
    vol = (0.6 - np.sqrt(xx * xx + yy * yy + zz * zz)).astype(np.float32)
    t0 = time.perf_counter()
    omc.marching_cubes(vol, 0.0)
    t_mc = time.perf_counter() - t0
    npts = (octree_resolution + 1) ** 3
    chunks = -(-npts // 16000)
    total = steps * (16 * t_double + 32 * t_single) + chunks * t_chunk + t_mc * npts / n ** 3
    details = dict(t_double_block_s=t_double, t_single_block_s=t_single, t_decode_chunk16000_s=t_chunk,
                   t_mc_s=t_mc, mc_grid=n, extrapolated_object_s=total,
                   sample=(f"{dit_reps}x(1 double + 1 single DiT block, B=2, L=4442) scaled x(16,32)x{steps} steps; "
                           f"{chunk_reps} geo-decoder chunks of 16000 queries scaled x{chunks}; C marching cubes on "
                           f"{n}^3 scaled x{npts / n ** 3:.1f}; fp32 torch CPU"),
                   sampled_cpu_seconds=dit_reps * (t_double + t_single) + chunk_reps * t_chunk + t_mc)
    return 1.0 / total, details


def time_vggt_sample(frames=2, reps=2):
    """BASELINE.json config 4 on the host: one frame-attention block + one global-attention block of the aggregator at
    the real geometry (S frames x 1374 tokens x 1024) averaged over `reps`, scaled to 24 DINOv2 + 24 frame + 24 global
    blocks (a DINOv2 block costs a frame block without q/k norm and RoPE), plus the reference's numpy back-projection
    (geometry.py:15-117) of S x 518 x 518 depth maps, timed in full.  The DPT / camera heads (~5 % of the FLOPs) are
    left out of the sample.  Returns (frames_per_second, details)."""
    import vggt_ref as V
    torch.manual_seed(0)
    C, P = 1024, 1374
    sd = {}
    for p in ("frame_blocks.0.", "global_blocks.0."):
        for n, shp in (("attn.qkv", (3 * C, C)), ("attn.proj", (C, C)), ("mlp.fc1", (4 * C, C)), ("mlp.fc2", (C, 4 * C))):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.randn(*shp) * 0.02, torch.zeros(shp[0])
        for n, d in (("norm1", C), ("norm2", C), ("attn.q_norm", 64), ("attn.k_norm", 64)):
            sd[p + n + ".weight"], sd[p + n + ".bias"] = torch.ones(d), torch.zeros(d)
        sd[p + "ls1.gamma"] = sd[p + "ls2.gamma"] = torch.full((C,), 0.1)
    x = torch.randn(frames, P, C)
    pos = torch.zeros(frames, P, 2, dtype=torch.long)
    t_frame = t_global = 0.0
    with torch.no_grad():
        for _ in range(reps):
            t0 = time.perf_counter()
            y = V.block(sd, "frame_blocks.0.", x, 16, 1e-5, pos, 100.0)
            t_frame += (time.perf_counter() - t0) / reps
            t0 = time.perf_counter()
            V.block(sd, "global_blocks.0.", y.view(1, frames * P, C), 16, 1e-5, pos.view(1, frames * P, 2), 100.0)
            t_global += (time.perf_counter() - t0) / reps
    rng = np.random.default_rng(0)
    depth = (rng.random((frames, 518, 518, 1)) + 0.5).astype(np.float32)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    E = np.tile(np.concatenate([q, rng.normal(size=(3, 1))], 1).astype(np.float32), (frames, 1, 1))
    K = np.tile(np.array([[500, 0, 259], [0, 500, 259], [0, 0, 1]], np.float32), (frames, 1, 1))
    t0 = time.perf_counter()
    R.unproject_depth_map_to_point_map(depth, E, K)
    t_unproject = time.perf_counter() - t0
    total = 48 * t_frame + 24 * t_global + t_unproject
    det = dict(t_frame_block_s=t_frame, t_global_block_s=t_global, t_unproject_s=t_unproject,
               extrapolated_scene_s=total, sampled_cpu_seconds=reps * (t_frame + t_global) + t_unproject,
               sample=(f"{reps}x(1 frame block + 1 global block, S={frames}, 1374 tokens, 1024 wide) scaled x(48, 24); "
                       f"numpy back-projection of {frames}x518x518 in full; fp32 torch CPU"))
    return frames / total, det
