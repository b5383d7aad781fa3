"""CPU: the marching-cubes oracle against topological / geometric invariants (config 1 of BASELINE.json:
64^3 synthetic sphere SDF -> marching cubes on CPU).  scikit-image is not available, so the oracle is
"parity unpinned" against it; these properties are what any correct Lewiner-style extraction satisfies."""
import numpy as np
import pytest

import mc as omc


def grid_axes(n, bound=1.01):
    ax = np.linspace(-bound, bound, n, dtype=np.float32)
    return np.meshgrid(ax, ax, ax, indexing="ij")


def sphere(n, r=0.6):
    x, y, z = grid_axes(n)
    return (r - np.sqrt(x * x + y * y + z * z)).astype(np.float32)


def torus(n, R=0.6, r=0.25):
    x, y, z = grid_axes(n)
    return (r - np.sqrt((np.sqrt(x * x + y * y) - R) ** 2 + z * z)).astype(np.float32)


def random_field(n, seed=0, coarse=8):
    import torch
    torch.manual_seed(seed)
    c = torch.randn(1, 1, coarse, coarse, coarse)
    f = torch.nn.functional.interpolate(c, size=(n, n, n), mode="trilinear", align_corners=True)
    return f[0, 0].numpy().astype(np.float32)


def edge_stats(faces):
    e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
    und = np.sort(e, axis=1)
    uniq, counts = np.unique(und, axis=0, return_counts=True)
    # directed edges must be unique and each undirected edge used once per direction (consistent orientation)
    d_uniq = np.unique(e, axis=0)
    return uniq, counts, len(d_uniq) == len(e)


def check_closed_manifold(verts, faces):
    assert faces.min() >= 0 and faces.max() < len(verts)
    uniq, counts, directed_unique = edge_stats(faces)
    assert (counts == 2).all(), "every edge must be shared by exactly two triangles (closed surface)"
    assert directed_unique, "orientation must be consistent"
    return len(verts) - len(uniq) + len(faces)  # Euler characteristic


def signed_volume(verts, faces):
    a, b, c = verts[faces[:, 0]].astype(np.float64), verts[faces[:, 1]].astype(np.float64), verts[faces[:, 2]].astype(np.float64)
    return np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0


def test_sphere_64_topology_and_geometry():
    n = 65
    vol = sphere(n)
    v, f, cases = omc.marching_cubes(vol, 0.0, return_cases=True)
    assert v.dtype == np.float32 and f.dtype == np.int32
    chi = check_closed_manifold(v, f)
    assert chi == 2
    # vertices lie on the sphere (index space -> world), within linear-interpolation error
    w = v / (n - 1) * 2.02 - 1.01
    rad = np.linalg.norm(w, axis=1)
    assert np.abs(rad - 0.6).max() < 2e-3
    # skimage 'descent' winding: right-hand normals point towards increasing values (into the object);
    # Hunyuan flips the faces afterwards (pipelines.py:102).  Inward normals <=> negative signed volume.
    assert signed_volume(w, f) < 0
    assert abs(abs(signed_volume(w, f)) - 4 / 3 * np.pi * 0.6 ** 3) < 5e-3
    # every used vertex is referenced, traversal order: first face of the first active cell uses vertex 0
    assert np.array_equal(np.unique(f), np.arange(len(v)))
    assert set(np.unique(cases)) <= set(range(15))


def test_torus_genus_one():
    v, f = omc.marching_cubes(torus(49), 0.0)
    assert check_closed_manifold(v, f) == 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_field_closed_and_ambiguous_cases(seed):
    n = 33
    vol = random_field(n, seed)
    # pad with a negative shell so the surface is closed inside the volume
    vol[0, :, :] = vol[-1, :, :] = vol[:, 0, :] = vol[:, -1, :] = vol[:, :, 0] = vol[:, :, -1] = -1.0
    v, f, cases = omc.marching_cubes(vol, 0.0, return_cases=True)
    chi = check_closed_manifold(v, f)
    assert chi % 2 == 0
    present = set(np.unique(cases))
    assert {3, 6, 7} & present, "the random field is meant to exercise ambiguous configurations"


def test_vertices_on_sign_changing_edges_and_formula():
    vol = random_field(17, 5)
    v, f = omc.marching_cubes(vol, 0.1)
    lvl = np.float32(0.1)
    frac = v - np.floor(v)
    on_edge = (frac > 0).sum(axis=1)
    centre = on_edge == 3   # cell-centre vertices of the >= 8-gon tilings
    assert (on_edge[~centre] <= 1).all()
    for p in v[~centre][:2000]:
        axis = int(np.argmax(p - np.floor(p) > 0)) if (p - np.floor(p) > 0).any() else 0
        lo = np.floor(p).astype(int)
        hi = lo.copy()
        hi[axis] = min(hi[axis] + 1, vol.shape[axis] - 1)
        a, b = float(vol[tuple(lo)]) - float(lvl), float(vol[tuple(hi)]) - float(lvl)
        if (p - np.floor(p) > 0).any():
            assert (a > 0) != (b > 0)
            eps = float(np.finfo(np.float32).eps)
            wa, wb = 1 / (eps + abs(a)), 1 / (eps + abs(b))
            expect = np.float32(lo[axis] + wb / (wa + wb))
            assert p[axis] == expect


def test_level_out_of_range_and_no_surface():
    vol = sphere(9)
    with pytest.raises(ValueError):
        omc.marching_cubes(vol, 10.0)
    flat = np.zeros((5, 5, 5), np.float32)
    with pytest.raises(RuntimeError):
        omc.marching_cubes(flat, 0.0)


def test_hunyuan_rescale_divides_by_grid_size():
    n = 17
    vol = sphere(n)
    v0, f0 = omc.marching_cubes(vol, 0.0)
    v1, f1 = omc.marching_cubes(vol, 0.0, bounds=[-1.01] * 3 + [1.01] * 3)
    expect = (v0 / [n, n, n] * np.array([2.02] * 3) + np.array([-1.01] * 3)).astype(np.float32)
    assert np.array_equal(v1, expect)
    assert np.array_equal(f0, f1)


def test_ragged_shapes():
    x, y, z = np.meshgrid(np.linspace(-1, 1, 7), np.linspace(-1, 1, 12), np.linspace(-1, 1, 9), indexing="ij")
    vol = (0.7 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)
    v, f = omc.marching_cubes(vol, 0.0)
    assert check_closed_manifold(v, f) == 2
    assert v[:, 0].max() <= 6 and v[:, 1].max() <= 11 and v[:, 2].max() <= 8
