#!/usr/bin/env python
"""Write scikit-image goldens for the marching-cubes parity tests.  Run this on ANY machine that has
`scikit-image>=0.24` (the pin of the reference, requirements.txt:17) and numpy -- it needs nothing else from this
repository -- and commit the files it writes under tests/golden/:

    python tools/dump_skimage_goldens.py            # -> tests/golden/skimage_mc_<name>.npz

This container and the GPU boxes have no scikit-image and no network, which is why the marching-cubes oracle is
"parity unpinned".  tests/test_mc_skimage_golden.py picks the files up as soon as they exist and reports, per volume,
which of {identical arrays, triangle set, vertex set, topology} agree with the reference call
    skimage.measure.marching_cubes(volume, level, method="lewiner")
(Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:69-73).  The volumes themselves are stored in the
files, so nothing depends on reproducing an RNG stream."""
import os

import numpy as np


def axes(n, bound=1.01):
    ax = np.linspace(-bound, bound, n, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def sphere(n, r=0.6):
    x, y, z = axes(n)
    return (r - np.sqrt(x * x + y * y + z * z)).astype(np.float32)


def torus(n, R=0.6, r=0.25):
    x, y, z = axes(n)
    return (r - np.sqrt((np.sqrt(x * x + y * y) - R) ** 2 + z * z)).astype(np.float32)


def smooth_field(n, seed, coarse=8):
    """Seeded coarse noise, trilinearly upsampled with numpy (ambiguous faces, smooth surface)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((coarse, coarse, coarse))
    t = np.linspace(0, coarse - 1, n)
    i0 = np.clip(np.floor(t).astype(int), 0, coarse - 2)
    w = t - i0
    for axis in range(3):
        a = np.take(c, i0, axis=axis)
        b = np.take(c, i0 + 1, axis=axis)
        shape = [1, 1, 1]
        shape[axis] = n
        ww = w.reshape(shape)
        c = a * (1 - ww) + b * ww
    return c.astype(np.float32)


def white_noise(shape, seed):
    """Every base case, every face-test sub-case, and the interior tests of cases 4 / 10 both ways."""
    vol = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    vol[0, :, :] = vol[-1, :, :] = vol[:, 0, :] = vol[:, -1, :] = vol[:, :, 0] = vol[:, :, -1] = -1.0
    return vol


def case_atlas(seed, per_case=6):
    """One isolated cell per draw: 256 cube indices x `per_case` magnitude draws, each cell alone in a 4^3 block of a
    larger volume (so a mismatch names one cube index).  Magnitudes are log-normal so that face and interior tests
    fall on both sides."""
    rng = np.random.default_rng(seed)
    n = 256 * per_case
    side = int(np.ceil(n ** (1 / 3)))
    vol = np.full((4 * side, 4 * side, 4 * side), -3.0, np.float32)
    corner = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
    k = 0
    for ci in range(1, 255):
        for _ in range(per_case):
            mag = np.exp(1.2 * rng.standard_normal(8)) * 0.5
            bz, by, bx = (k // (side * side)) * 4, ((k // side) % side) * 4, (k % side) * 4
            for i, (x, y, z) in enumerate(corner):
                vol[bz + 1 + z, by + 1 + y, bx + 1 + x] = mag[i] if (ci >> i) & 1 else -mag[i]
            k += 1
    return vol


VOLUMES = [
    ("sphere33", lambda: sphere(33), 0.0), ("sphere65", lambda: sphere(65), 0.0),
    ("torus33", lambda: torus(33), 0.0), ("torus65", lambda: torus(65), 0.0),
    ("smooth33_s0", lambda: smooth_field(33, 0), 0.0), ("smooth65_s1", lambda: smooth_field(65, 1), 0.1),
    ("noise17_s0", lambda: white_noise((17, 17, 17), 0), 0.0), ("noise_ragged_s1", lambda: white_noise((12, 15, 19), 1), 0.0),
    ("atlas_s0", lambda: case_atlas(0), 0.0), ("atlas_s1", lambda: case_atlas(1), 0.0),
]


def main():
    import skimage
    from skimage import measure
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, make, level in VOLUMES:
        vol = make()
        verts, faces, _, _ = measure.marching_cubes(vol, level, method="lewiner")
        path = os.path.join(out_dir, f"skimage_mc_{name}.npz")
        np.savez_compressed(path, volume=vol, level=np.float32(level), verts=verts.astype(np.float32),
                            faces=np.ascontiguousarray(faces).astype(np.int32), skimage_version=skimage.__version__)
        print(f"{path}: V={len(verts)} F={len(faces)} (scikit-image {skimage.__version__})")


if __name__ == "__main__":
    main()
