#!/usr/bin/env python
"""Stage 4 (`python run.py -p 4`) -- model part of src/camera_and_pointcloud/minimal_demo_vggt.py on r3g.

Covered: load_and_preprocess_images_square at 1024 (PIL bicubic),
run_VGGT = bilinear resize to 518 -> aggregator (r3g kernels) -> camera head -> pose_encoding_to_extri_intri ->
DPT depth head -> unproject_depth_map_to_point_map (r3g kernel, float64) -> confidence mask ->
randomly_limit_trues; outputs `points.ply` (+ `vggt_raw.npz`: extrinsic, intrinsic, depth, conf) under
`config["output_vggt"]`.

NOT covered here (section 8f row 3, "next"): the COLMAP export through pycolmap
(minimal_demo_vggt.py:487-578) and export_vggt_data's coordinate fixes (:76-262) that produce `camera.npz` /
`scene_vggt.ply` -- pycolmap is not installed in this image and those are wire-format writers on the tail.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F
import yaml
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "3d-re-gen_b200"))

from r3g import ops  # noqa: E402
from r3g.vggt_heads import VGGT, pose_encoding_to_extri_intri  # noqa: E402


def load_and_preprocess_images_square(paths, target_size=1024):
    """vggt/vggt/utils/load_fn.py:13-94."""
    if len(paths) == 0:
        raise ValueError("At least 1 image is required")
    images, coords = [], []
    for path in paths:
        img = Image.open(path)
        if img.mode == "RGBA":
            img = Image.alpha_composite(Image.new("RGBA", img.size, (255, 255, 255, 255)), img)
        img = img.convert("RGB")
        w, h = img.size
        side = max(w, h)
        left, top = (side - w) // 2, (side - h) // 2
        s = target_size / side
        coords.append(np.array([left * s, top * s, (left + w) * s, (top + h) * s, w, h]))
        sq = Image.new("RGB", (side, side), (0, 0, 0))
        sq.paste(img, (left, top))
        sq = sq.resize((target_size, target_size), Image.Resampling.BICUBIC)
        images.append(torch.from_numpy(np.asarray(sq).astype(np.float32) / 255.0).permute(2, 0, 1))
    return torch.stack(images), torch.from_numpy(np.array(coords)).float()


def randomly_limit_trues(mask, max_trues):
    """vggt/vggt/utils/helper.py:10-30 (np.random.choice => seeded by np.random.seed)."""
    idx = np.flatnonzero(mask)
    if idx.size <= max_trues:
        return mask
    keep = np.random.choice(idx, size=max_trues, replace=False)
    out = np.zeros(mask.size, dtype=bool)
    out[keep] = True
    return out.reshape(mask.shape)


def run_VGGT(model, images, resolution=518):
    """minimal_demo_vggt.py:295-321; returns numpy extrinsic [S,3,4], intrinsic [S,3,3], and the DEVICE depth/conf."""
    images = F.interpolate(images, size=(resolution, resolution), mode="bilinear", align_corners=False)[None]
    tokens, ps_idx = model.aggregator(images)
    pose_enc = model.camera_head(tokens)[-1]
    extrinsic, intrinsic = pose_encoding_to_extri_intri(pose_enc, images.shape[-2:])
    depth, conf = model.depth_head(tokens, images, ps_idx)
    return extrinsic[0].cpu().numpy(), intrinsic[0].cpu().numpy(), depth[0], conf[0]


def write_ply(path, pts, rgb):
    with open(path, "wb") as fh:
        fh.write((f"ply\nformat binary_little_endian 1.0\nelement vertex {len(pts)}\nproperty float x\nproperty float y\n"
                  "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n").encode())
        rec = np.empty(len(pts), dtype=[("p", "<f4", 3), ("c", "u1", 3)])
        rec["p"], rec["c"] = pts, rgb
        fh.write(rec.tobytes())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="../src/config.yaml")
    ap.add_argument("--checkpoint", default=None, help="VGGT-1B state_dict (.pt); random weights if absent")
    args = ap.parse_args()
    cfg = yaml.safe_load(open(args.config))
    seed = cfg.get("seed", 42)
    np.random.seed(seed)
    torch.manual_seed(seed)
    out_dir = cfg["output_vggt"]
    os.makedirs(out_dir, exist_ok=True)
    paths = [cfg["image_url"]]
    empty = "../output/findings/banana/inpaint_nanoBanana/empty_room.png"
    if os.path.exists(empty):
        paths.append(empty)
    model = VGGT()
    if args.checkpoint and os.path.exists(args.checkpoint):
        model.load_state_dict(torch.load(args.checkpoint, map_location="cpu", weights_only=True))
    else:
        raise SystemExit("no VGGT checkpoint reachable (no network): pass --checkpoint, or see tools/bench_vggt.py "
                         "for the random-weight timing run")
    images, _ = load_and_preprocess_images_square(paths, cfg.get("img_load_resolution", 1024))
    images = images.cuda()
    extrinsic, intrinsic, depth, conf = run_VGGT(model, images)
    pts = ops.unproject(depth[..., 0].contiguous(), extrinsic, intrinsic, torch.float64).cpu().numpy()
    conf_np = conf.cpu().numpy()
    mask = randomly_limit_trues(conf_np >= cfg.get("conf_thres_value", 1.5), cfg.get("max_points_for_colmap", 100000))
    rgb = (F.interpolate(images, size=depth.shape[1:3], mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
           .cpu().numpy() * 255).astype(np.uint8)
    write_ply(os.path.join(out_dir, "points.ply"), pts[mask].astype(np.float32), rgb[mask])
    np.savez(os.path.join(out_dir, "vggt_raw.npz"), extrinsic=extrinsic, intrinsic=intrinsic,
             depth=depth.cpu().numpy(), conf=conf_np)
    print(f"wrote {int(mask.sum())} points to {out_dir}")


if __name__ == "__main__":
    main()
