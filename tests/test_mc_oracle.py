"""CPU: the marching-cubes oracle (oracle/mc_oracle.c, table driven) against
  * topological / geometric invariants (config 1 of BASELINE.json: 64^3 synthetic sphere SDF -> marching cubes on CPU),
  * a SECOND, independent implementation (oracle/mc_tracer.py: every cell traced from the rules, no generated tables,
    no ownership predicate) -- bit for bit, order included,
  * an independent numpy check of the topology on every grid face (iso-segments predicted from the four corner signs
    and Lewiner's face test).
scikit-image is not available, so the oracle stays "parity unpinned" against it (tests/test_mc_skimage_golden.py
activates when goldens exist); these properties are what any correct Lewiner-style extraction satisfies."""
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
            eps = float(np.spacing(1.0))      # scikit-image's FLT_EPSILON (include/r3g_mc_tables.h: R3G_MC_EPS)
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


def noise(shape, seed):
    vol = np.random.default_rng(seed).standard_normal(shape).astype(np.float32)
    vol[0, :, :] = vol[-1, :, :] = vol[:, 0, :] = vol[:, -1, :] = vol[:, :, 0] = vol[:, :, -1] = -1.0   # closed surface
    return vol


@pytest.mark.parametrize("seed,shape", [(0, (12, 13, 14)), (1, (14, 12, 13)), (2, (9, 9, 9))])
def test_table_oracle_equals_independent_tracer_on_white_noise(seed, shape):
    """White noise holds every base case and every face-test sub-case, and the interior test fires both ways
    (4.1.1 / 4.1.2, 10.1.1 / 10.1.2)."""
    import mc_tracer
    vol = noise(shape, seed)
    v, f, cases = omc.marching_cubes(vol, 0.0, return_cases=True)
    tv, tf, sub = mc_tracer.marching_cubes(vol, 0.0, return_subcases=True)
    assert np.array_equal(v, tv) and np.array_equal(f, tf)
    assert check_closed_manifold(v, f) % 2 == 0
    if seed == 0:
        assert set(np.unique(cases)) == set(range(15))
        kinds = {(k[0], k[2]) for k in sub}
        assert (4, "tunnel") in kinds and (4, "tested-separate") in kinds
        assert {(b, "tested-separate") for b in (6, 7, 10, 12)} <= kinds


def test_tracer_equals_oracle_on_smooth_fields_and_level_shift():
    import mc_tracer
    for vol, lvl in ((random_field(15, 3), 0.1), (torus(17), 0.0), (sphere(13), -0.05)):
        v, f = omc.marching_cubes(vol, lvl)
        tv, tf = mc_tracer.marching_cubes(vol, lvl)
        assert np.array_equal(v, tv) and np.array_equal(f, tf)


def test_tunnel_cells_of_case_10():
    """Two strong opposite edges joined through a weak interior: Lewiner's test_interior picks 10.1.2."""
    import mc_tracer
    found = 0
    rng = np.random.default_rng(7)
    for _ in range(4000):
        c = rng.standard_normal(8) * np.exp(1.5 * rng.standard_normal(8))
        c = np.abs(c) * np.array([1, 1, -1, -1, -1, -1, 1, 1])      # corners 0,1 and 6,7 positive: base case 10
        vol = np.full((4, 4, 4), -5.0, np.float32)
        for i, (x, y, z) in enumerate([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]):
            vol[1 + z, 1 + y, 1 + x] = c[i]
        tv, tf, sub = mc_tracer.marching_cubes(vol, 0.0, return_subcases=True)
        if any(k[0] == 10 and k[2] == "tunnel" for k in sub):
            found += 1
            v, f = omc.marching_cubes(vol, 0.0)
            assert np.array_equal(v, tv) and np.array_equal(f, tf)
            assert check_closed_manifold(v, f) == 2          # ONE component of genus 0 through the tunnel
    assert found >= 5


def test_face_topology_against_an_independent_numpy_prediction():
    """On every grid face (the square between two cells) the mesh must carry exactly the iso-segments that the four
    corner signs and Lewiner's face test predict: computed here with numpy from the volume alone, no tables."""
    vol = noise((11, 12, 10), 5)
    v, f = omc.marching_cubes(vol, 0.0)
    eps = float(np.spacing(1.0))
    # key of a vertex: the grid edge it sits on = (axis, lower grid point); centre vertices have three fractional parts
    frac = v - np.floor(v)
    nfrac = (frac > 0).sum(1)
    lo = np.floor(v).astype(int)
    axis = np.argmax(frac > 0, axis=1)
    key = {}
    for i in range(len(v)):
        if nfrac[i] <= 1:
            key[i] = (int(axis[i]) if nfrac[i] == 1 else -1, tuple(lo[i]))
    und = set()
    for a, b, c in f:
        for p, q in ((a, b), (b, c), (c, a)):
            und.add((min(p, q), max(p, q)))
    by_edge = {}
    for i, k in key.items():
        by_edge.setdefault(k, i)
    n = vol.shape
    checked = extra = 0
    for normal in range(3):
        u_ax, w_ax = [a for a in range(3) if a != normal]
        for p0 in np.ndindex(*[n[a] - (0 if a == normal else 1) for a in range(3)]):
            p = np.array(p0)
            if p[normal] in (0, n[normal] - 1):
                continue                                # faces on the volume boundary have a cell on one side only
            def val(du, dw):
                q = p.copy(); q[u_ax] += du; q[w_ax] += dw
                return float(vol[tuple(q)])
            c00, c10, c11, c01 = val(0, 0), val(1, 0), val(1, 1), val(0, 1)       # cyclic A, B, C, D
            s = [c > 0 for c in (c00, c10, c11, c01)]
            if sum(s) in (0, 4):
                continue
            def edge_vertex(k):                          # vertex on side k of the square (A-B, B-C, C-D, D-A)
                q = p.copy()
                if k == 0: ax = u_ax
                elif k == 1: q[u_ax] += 1; ax = w_ax
                elif k == 2: q[w_ax] += 1; ax = u_ax
                else: ax = w_ax
                return by_edge[(ax, tuple(q))]
            crossing = [k for k in range(4) if s[k] != s[(k + 1) % 4]]
            if len(crossing) == 2:
                expect = {tuple(sorted((edge_vertex(crossing[0]), edge_vertex(crossing[1]))))}
            else:
                acbd = c00 * c11 - c10 * c01
                pos_joined = True if abs(acbd) < eps else ((acbd >= 0) == (c00 > 0))
                # corners cut off: the negative ones if the positive pair is joined, else the positive ones
                cut = [k for k in range(4) if s[k] != pos_joined]
                expect = {tuple(sorted((edge_vertex((k - 1) % 4), edge_vertex(k)))) for k in cut}
            on_face = {edge_vertex(k) for k in crossing}
            have = {e for e in und if e[0] in on_face and e[1] in on_face}
            assert expect <= have, (p0, normal)
            extra += len(have - expect)                  # rungs of tunnel tilings may lie in a face (7.4.2 only)
            checked += 1
    assert checked > 500 and extra == 0
