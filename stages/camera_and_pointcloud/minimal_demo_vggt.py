#!/usr/bin/env python
"""Stage 4 (`python run.py -p 4`) -- src/camera_and_pointcloud/minimal_demo_vggt.py on r3g (feed-forward branch,
`use_ba: false`, the configuration 3D-RE-GEN ships).

Rows v1-v7 of SURVEY.md section 8a: load_and_preprocess_images_square at 1024 (PIL bicubic), run_VGGT = bilinear resize
to 518 -> aggregator (r3g kernels) -> camera head -> pose_encoding_to_extri_intri -> DPT depth head ->
unproject_depth_map_to_point_map (r3g kernel, float64).
Row v8, the tail (minimal_demo_vggt.py:458-578 and export_vggt_data :76-262): confidence mask + randomly_limit_trues,
the COLMAP sparse model (cameras/images/points3D.bin, written by colmap_io.py without pycolmap's per-point Python loop),
image_list.txt, points_merged.ply / points.ply / points_emptyRoom(_pre).ply, then camera.npz, camera_emptyRoom.npz and
the scene cloud (`config["vggt_cloud"]`) with the OpenCV -> Blender -> PyTorch3D coordinate fixes.
Not covered: the bundle-adjustment branch (`use_ba: true`: VGGSfM tracker + pycolmap BA, third-party solvers).
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

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import colmap_io  # noqa: E402

try:  # the model part needs the CUDA library; the tail (below) is host code and importable without it
    from r3g import ops  # noqa: E402
    from r3g.vggt_heads import VGGT, pose_encoding_to_extri_intri  # noqa: E402
except Exception as _e:  # pragma: no cover - reported when main() is reached
    ops = VGGT = pose_encoding_to_extri_intri = None
    _IMPORT_ERROR = _e


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


def write_ply(path, pts, rgb=None):
    """Binary little-endian PLY point cloud (float xyz [+ uchar rgb]); the reference exports through
    trimesh.PointCloud(...).export, downstream reads `.vertices` only."""
    pts = np.asarray(pts, np.float32).reshape(-1, 3)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as fh:
        hdr = f"ply\nformat binary_little_endian 1.0\nelement vertex {len(pts)}\nproperty float x\nproperty float y\nproperty float z\n"
        if rgb is not None:
            hdr += "property uchar red\nproperty uchar green\nproperty uchar blue\n"
        fh.write((hdr + "end_header\n").encode())
        if rgb is None:
            fh.write(pts.astype("<f4").tobytes())
        else:
            rec = np.empty(len(pts), dtype=[("p", "<f4", 3), ("c", "u1", 3)])
            rec["p"], rec["c"] = pts, np.asarray(rgb, np.uint8).reshape(-1, 3)
            fh.write(rec.tobytes())


def read_ply_vertices(path):
    with open(path, "rb") as fh:
        n, props = 0, []
        while True:
            ln = fh.readline().decode().strip()
            if ln.startswith("element vertex"):
                n = int(ln.split()[-1])
            elif ln.startswith("property"):
                props.append(ln.split()[1:])
            elif ln == "end_header":
                break
        dt = np.dtype([(name, {"float": "<f4", "uchar": "u1"}[ty]) for ty, name in props])
        rec = np.frombuffer(fh.read(n * dt.itemsize), dt)
    return np.stack([rec["x"], rec["y"], rec["z"]], axis=1)


def create_pixel_coordinate_grid(num_frames, height, width):
    """vggt/vggt/utils/helper.py:33-60: [S,H,W,3] float32 of (x, y, frame)."""
    y, x = np.indices((height, width), dtype=np.float32)
    f = np.arange(num_frames, dtype=np.float32)[:, None, None]
    return np.stack((np.broadcast_to(x[None], (num_frames, height, width)),
                     np.broadcast_to(y[None], (num_frames, height, width)),
                     np.broadcast_to(f, (num_frames, height, width))), axis=-1)


def B2P(B):
    """src/utils/global_utils.py:835-844: Blender 4x4 -> PyTorch3D (R, T)."""
    r1 = np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], dtype=np.float64)
    r2 = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, -1]], dtype=np.float64)
    tt = np.array([[-1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float64)
    R = r1 @ B[:3, :3] @ r2
    return R, tt @ B[:3, 3] @ R


def write_sparse_model(out_dir, points_3d, conf, images_518, extrinsic, intrinsic, image_paths, original_coords, cfg,
                       resolution=518):
    """minimal_demo_vggt.py:458-578 (feed-forward branch).  points_3d [S,H,W,3], conf [S,H,W], images_518 [S,3,H,W] in
    [0,1] (torch or numpy); returns the path of the main cloud."""
    points_3d = np.asarray(points_3d)
    S, H, W, _ = points_3d.shape
    rgb = (np.asarray(images_518) * 255).astype(np.uint8).transpose(0, 2, 3, 1)
    xyf = create_pixel_coordinate_grid(S, H, W)
    mask = randomly_limit_trues(np.asarray(conf) >= cfg.get("conf_thres_value", 5.0), cfg.get("max_points_for_colmap", 100000))
    rc = colmap_io.build_reconstruction_wo_track(points_3d[mask], xyf[mask], rgb[mask], extrinsic, intrinsic,
                                                 np.array([resolution, resolution]), shared_camera=False,
                                                 camera_type="PINHOLE")
    colmap_io.rename_and_rescale(rc, image_paths, np.asarray(original_coords), img_size=resolution,
                                 shift_point2d_to_original_res=True, shared_camera=False)
    rc.write(out_dir)
    with open(os.path.join(out_dir, "image_list.txt"), "w") as fh:
        for p in image_paths:
            fh.write(str(p) + "\n")
    write_ply(os.path.join(out_dir, "points_merged.ply"), points_3d[mask], rgb[mask])
    main_ply = os.path.join(out_dir, "points.ply")
    pts0, rgb0 = points_3d[0][mask[0]], rgb[0][mask[0]]
    write_ply(main_ply, pts0, rgb0)
    if S >= 2:   # the inpainted empty room: saved raw and with its bounding box scaled to the main cloud's
        pts1, rgb1 = points_3d[1][mask[1]], rgb[1][mask[1]]
        write_ply(os.path.join(out_dir, "points_emptyRoom_pre.ply"), pts1, rgb1)
        fit = pts1
        if pts1.size and pts0.size:
            src_ext, tgt_ext = pts1.max(0) - pts1.min(0), pts0.max(0) - pts0.min(0)
            scale = np.divide(tgt_ext, src_ext, out=np.ones_like(tgt_ext), where=src_ext > 1e-6)
            c = pts1.mean(0)
            fit = (pts1 - c) * scale + c
        write_ply(os.path.join(out_dir, "points_emptyRoom.ply"), fit, rgb1)
    return main_ply


def export_vggt_data(cfg):
    """minimal_demo_vggt.py:76-262 on colmap_io.Reconstruction: camera.npz (+ camera_emptyRoom.npz) and the scene cloud."""
    rdir = cfg.get("output_vggt", "../output/vggt/sparse")
    rc = colmap_io.Reconstruction.read(rdir)
    names = []
    lst = os.path.join(rdir, "image_list.txt")
    if os.path.exists(lst):
        names = [ln.strip() for ln in open(lst) if ln.strip()]

    def find(name):
        for iid in sorted(rc.images):
            n = rc.images[iid]["name"]
            if n == name or os.path.basename(n) == os.path.basename(name):
                return iid
        return None

    main_id = find(names[0]) if names else None
    if main_id is None:
        main_id = sorted(rc.images)[0]
    R_fix = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0]], dtype=np.float32)

    def camera_record(iid):
        T = rc.cam_from_world(iid).astype(np.float32)
        ext = np.eye(4, dtype=np.float32)
        ext[:3, :3] = R_fix @ T[:, :3]
        ext[:3, 3] = R_fix @ T[:, 3]
        cam = rc.cameras[rc.images[iid]["camera_id"]]
        fx, fy = cam["params"][0], cam["params"][1]
        focal = float((fx + fy) / 2.0)
        return {"extrinsic": ext, "focal": np.float32(focal),
                "image_size": np.array([cam["width"], cam["height"]], dtype=np.int32),
                "camera_angle_x": np.float32(2.0 * np.arctan(cam["width"] / (2.0 * focal)))}

    points = read_ply_vertices(os.path.join(rdir, "points.ply"))
    rec = camera_record(main_id)
    R_p3d, T_p3d = B2P(rec["extrinsic"])
    pts = (points @ R_fix.T) @ R_p3d.T + T_p3d
    pts[:, 1] *= -1
    pts *= cfg.get("vggt_scene_scale", 5.0)
    path = os.path.abspath(cfg["camera"])
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, **rec)
    if len(names) >= 2 and find(names[1]) is not None:
        np.savez(os.path.join(os.path.dirname(path), "camera_emptyRoom.npz"), **camera_record(find(names[1])))
    write_ply(os.path.abspath(cfg["vggt_cloud"]), pts)
    return rec, pts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="../src/config.yaml")
    ap.add_argument("--checkpoint", default=None, help="VGGT-1B state_dict (.pt); there is no network to fetch it")
    args = ap.parse_args()
    if VGGT is None:
        raise SystemExit(f"r3g is not usable here: {_IMPORT_ERROR}")
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
    if not (args.checkpoint and os.path.exists(args.checkpoint)):
        raise SystemExit("no VGGT checkpoint reachable (no network): pass --checkpoint, or see tools/bench_vggt.py "
                         "for the random-weight timing run")
    model = VGGT()
    model.load_state_dict(torch.load(args.checkpoint, map_location="cpu", weights_only=True))
    images, original_coords = load_and_preprocess_images_square(paths, cfg.get("img_load_resolution", 1024))
    images = images.cuda()
    extrinsic, intrinsic, depth, conf = run_VGGT(model, images)
    pts = ops.unproject(depth[..., 0].contiguous(), extrinsic, intrinsic, torch.float64).cpu().numpy()
    img518 = F.interpolate(images, size=depth.shape[1:3], mode="bilinear", align_corners=False).cpu().numpy()
    write_sparse_model(out_dir, pts, conf.cpu().numpy(), img518, extrinsic, intrinsic, paths, original_coords.numpy(), cfg,
                       resolution=depth.shape[1])
    export_vggt_data(cfg)
    print(f"stage 4 written under {out_dir}")


if __name__ == "__main__":
    main()
