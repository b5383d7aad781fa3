"""COLMAP sparse-model files (cameras.bin / images.bin / points3D.bin) without pycolmap.

The reference builds a `pycolmap.Reconstruction` point by point in Python
(vggt/vggt/dependency/np_to_pycolmap.py:201-290, one `add_point3D` + one `Point2D` + one track element per point, 100 000
points per frame) and calls `reconstruction.write(dir)` (src/camera_and_pointcloud/minimal_demo_vggt.py:512).  This module
writes the same three files straight from the numpy arrays (vectorised, structured dtypes) and reads them back, in
COLMAP's documented binary layout (little endian):

  cameras.bin   u64 n | per camera: u32 id, i32 model, u64 width, u64 height, f64 params[k]
  images.bin    u64 n | per image:  u32 id, f64 qvec[4] (w,x,y,z), f64 tvec[3], u32 camera_id, name\\0,
                                    u64 n2d, n2d x (f64 x, f64 y, i64 point3D_id)
  points3D.bin  u64 n | per point:  u64 id, f64 xyz[3], u8 rgb[3], f64 error, u64 track_len,
                                    track_len x (u32 image_id, u32 point2D_idx)

pycolmap is not installed in this image, so these writers are checked by their own reader and by the format's
invariants (tests/test_stage4_tail.py), not against pycolmap: parity unpinned for this row.
"""
import os
import struct

import numpy as np

CAMERA_MODELS = {"SIMPLE_PINHOLE": (0, 3), "PINHOLE": (1, 4)}
_MODEL_BY_ID = {v[0]: (k, v[1]) for k, v in CAMERA_MODELS.items()}


def rotmat_to_qvec(R):
    """Rotation matrix -> unit quaternion (w, x, y, z), Eigen's `Quaternion(Matrix3)` branch structure (what
    pycolmap.Rotation3d(R) runs): trace > 0 -> w first, else the largest diagonal element leads."""
    R = np.asarray(R, dtype=np.float64)
    t = np.trace(R)
    q = np.empty(4)
    if t > 0.0:
        s = np.sqrt(t + 1.0)
        q[0] = 0.5 * s
        s = 0.5 / s
        q[1] = (R[2, 1] - R[1, 2]) * s
        q[2] = (R[0, 2] - R[2, 0]) * s
        q[3] = (R[1, 0] - R[0, 1]) * s
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[1 + i] = 0.5 * s
        s = 0.5 / s
        q[0] = (R[k, j] - R[j, k]) * s
        q[1 + j] = (R[j, i] + R[i, j]) * s
        q[1 + k] = (R[k, i] + R[i, k]) * s
    return q / np.linalg.norm(q)


def qvec_to_rotmat(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


class Reconstruction:
    """Arrays-of-structs view of a COLMAP sparse model (only what stage 4 writes and stage 4b reads)."""

    def __init__(self):
        self.cameras = {}    # id -> dict(model, width, height, params f64[k])
        self.images = {}     # id -> dict(name, camera_id, qvec f64[4], tvec f64[3], xys f64[n,2], point3D_ids i64[n])
        self.points_xyz = np.zeros((0, 3))       # row i has point3D id i + 1
        self.points_rgb = np.zeros((0, 3), np.uint8)
        self.points_error = np.zeros(0)
        self.track_image = np.zeros(0, np.uint32)    # one track element per point in the no-track form
        self.track_p2d = np.zeros(0, np.uint32)

    def cam_from_world(self, image_id):
        im = self.images[image_id]
        return np.concatenate([qvec_to_rotmat(im["qvec"]), im["tvec"][:, None]], axis=1)   # 3x4

    # ------------------------------------------------------------------------------------------ writers
    def write(self, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "cameras.bin"), "wb") as fh:
            fh.write(struct.pack("<Q", len(self.cameras)))
            for cid in sorted(self.cameras):
                c = self.cameras[cid]
                mid, npar = CAMERA_MODELS[c["model"]]
                assert len(c["params"]) == npar
                fh.write(struct.pack("<IiQQ", cid, mid, int(c["width"]), int(c["height"])))
                fh.write(np.asarray(c["params"], "<f8").tobytes())
        with open(os.path.join(out_dir, "images.bin"), "wb") as fh:
            fh.write(struct.pack("<Q", len(self.images)))
            for iid in sorted(self.images):
                im = self.images[iid]
                fh.write(struct.pack("<I", iid))
                fh.write(np.asarray(im["qvec"], "<f8").tobytes())
                fh.write(np.asarray(im["tvec"], "<f8").tobytes())
                fh.write(struct.pack("<I", im["camera_id"]))
                fh.write(im["name"].encode("utf-8") + b"\x00")
                n = len(im["xys"])
                fh.write(struct.pack("<Q", n))
                rec = np.empty(n, dtype=[("xy", "<f8", 2), ("pid", "<i8")])
                rec["xy"], rec["pid"] = im["xys"], im["point3D_ids"]
                fh.write(rec.tobytes())
        with open(os.path.join(out_dir, "points3D.bin"), "wb") as fh:
            n = len(self.points_xyz)
            fh.write(struct.pack("<Q", n))
            rec = np.empty(n, dtype=[("id", "<u8"), ("xyz", "<f8", 3), ("rgb", "u1", 3), ("err", "<f8"),
                                     ("tl", "<u8"), ("img", "<u4"), ("p2d", "<u4")])
            rec["id"] = np.arange(1, n + 1, dtype=np.uint64)
            rec["xyz"], rec["rgb"], rec["err"] = self.points_xyz, self.points_rgb, self.points_error
            rec["tl"], rec["img"], rec["p2d"] = 1, self.track_image, self.track_p2d
            fh.write(rec.tobytes())

    # ------------------------------------------------------------------------------------------ reader
    @classmethod
    def read(cls, in_dir):
        rc = cls()
        with open(os.path.join(in_dir, "cameras.bin"), "rb") as fh:
            (n,) = struct.unpack("<Q", fh.read(8))
            for _ in range(n):
                cid, mid, w, h = struct.unpack("<IiQQ", fh.read(24))
                name, npar = _MODEL_BY_ID[mid]
                rc.cameras[cid] = dict(model=name, width=w, height=h, params=np.frombuffer(fh.read(8 * npar), "<f8").copy())
        with open(os.path.join(in_dir, "images.bin"), "rb") as fh:
            (n,) = struct.unpack("<Q", fh.read(8))
            for _ in range(n):
                (iid,) = struct.unpack("<I", fh.read(4))
                q = np.frombuffer(fh.read(32), "<f8").copy()
                t = np.frombuffer(fh.read(24), "<f8").copy()
                (cid,) = struct.unpack("<I", fh.read(4))
                name = b""
                while True:
                    ch = fh.read(1)
                    if ch in (b"\x00", b""):
                        break
                    name += ch
                (n2,) = struct.unpack("<Q", fh.read(8))
                rec = np.frombuffer(fh.read(24 * n2), dtype=[("xy", "<f8", 2), ("pid", "<i8")])
                rc.images[iid] = dict(name=name.decode("utf-8"), camera_id=cid, qvec=q, tvec=t, xys=rec["xy"].copy(),
                                      point3D_ids=rec["pid"].copy())
        with open(os.path.join(in_dir, "points3D.bin"), "rb") as fh:
            (n,) = struct.unpack("<Q", fh.read(8))
            xyz, rgb, err, ti, tp = [], [], [], [], []
            for _ in range(n):   # general reader (variable track lengths); the writer above always emits length 1
                _pid, = struct.unpack("<Q", fh.read(8))
                xyz.append(np.frombuffer(fh.read(24), "<f8"))
                rgb.append(np.frombuffer(fh.read(3), "u1"))
                e, tl = struct.unpack("<dQ", fh.read(16))
                tr = np.frombuffer(fh.read(8 * tl), "<u4").reshape(tl, 2)
                err.append(e)
                ti.append(tr[0, 0] if tl else 0)
                tp.append(tr[0, 1] if tl else 0)
            rc.points_xyz = np.array(xyz).reshape(-1, 3)
            rc.points_rgb = np.array(rgb, np.uint8).reshape(-1, 3)
            rc.points_error = np.array(err)
            rc.track_image, rc.track_p2d = np.array(ti, np.uint32), np.array(tp, np.uint32)
        return rc


def build_reconstruction_wo_track(points3d, points_xyf, points_rgb, extrinsics, intrinsics, image_size,
                                  shared_camera=False, camera_type="PINHOLE"):
    """np_to_pycolmap.py:201-290 without its per-point loop.  points3d [P,3], points_xyf [P,3] (x, y, frame), points_rgb
    [P,3] uint8, extrinsics [N,3,4] (cam from world), intrinsics [N,3,3], image_size (w, h).  Point i gets id i + 1 and a
    one-element track (frame + 1, rank of the point among its frame's points); every frame gets image id frame + 1, named
    `image_{id}`, and -- unless shared_camera -- its own camera with the same id."""
    if camera_type not in CAMERA_MODELS:
        raise ValueError(f"Camera type {camera_type} is not supported yet")
    rc = Reconstruction()
    P = len(points3d)
    frame = points_xyf[:, 2].astype(np.int32)
    rc.points_xyz = np.asarray(points3d, np.float64).reshape(P, 3)
    rc.points_rgb = np.asarray(points_rgb, np.uint8).reshape(P, 3)
    rc.points_error = np.full(P, -1.0)      # pycolmap.Point3D's default error
    rc.track_image = (frame + 1).astype(np.uint32)
    rc.track_p2d = np.zeros(P, np.uint32)
    cam_id = None
    for f in range(len(extrinsics)):
        if cam_id is None or not shared_camera:
            K = intrinsics[f]
            params = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]]) if camera_type == "PINHOLE" else \
                np.array([(K[0, 0] + K[1, 1]) / 2, K[0, 2], K[1, 2]])
            cam_id = f + 1
            rc.cameras[cam_id] = dict(model=camera_type, width=int(image_size[0]), height=int(image_size[1]),
                                      params=params.astype(np.float64))
        idx = np.nonzero(frame == f)[0]
        rc.track_p2d[idx] = np.arange(len(idx), dtype=np.uint32)
        rc.images[f + 1] = dict(name=f"image_{f + 1}", camera_id=cam_id, qvec=rotmat_to_qvec(extrinsics[f][:3, :3]),
                                tvec=np.asarray(extrinsics[f][:3, 3], np.float64).copy(),
                                xys=np.asarray(points_xyf[idx, :2], np.float64), point3D_ids=(idx + 1).astype(np.int64))
    return rc


def rename_and_rescale(rc, image_paths, original_coords, img_size, shift_point2d_to_original_res=False,
                       shared_camera=False):
    """minimal_demo_vggt.py:325-363: names from the input paths, intrinsics rescaled from the padded `img_size` square to
    the original resolution (principal point = image centre), 2-D points shifted/scaled to original pixels."""
    rescale = True
    for iid in sorted(rc.images):
        im = rc.images[iid]
        cam = rc.cameras[im["camera_id"]]
        im["name"] = str(image_paths[iid - 1])
        real = np.asarray(original_coords[iid - 1, -2:], np.float64)
        ratio = max(real) / img_size
        if rescale:
            p = cam["params"] * ratio
            p[-2:] = real / 2
            cam["params"] = p
            cam["width"], cam["height"] = int(real[0]), int(real[1])
        if shift_point2d_to_original_res:
            im["xys"] = (im["xys"] - np.asarray(original_coords[iid - 1, :2], np.float64)) * ratio
        if shared_camera:
            rescale = False
    return rc
