"""ORACLE -- test infrastructure only.  Order-independent comparison of two triangle meshes of the same iso-surface.

Marching-cubes outputs can agree on geometry and topology and still differ in vertex order, triangle order, the
rotation of a triangle's three indices and the diagonal chosen inside a polygon.  `compare` separates those levels, so a
scikit-image golden (tests/golden/skimage_mc_*.npz, written by tools/dump_skimage_goldens.py on a machine that has
scikit-image) can pin each of {vertex set, topology, triangle set, order} independently of the others."""
import numpy as np


def _order(v):
    """Lexicographic order of the vertices (exact, on the float32 values)."""
    return np.lexsort((v[:, 2], v[:, 1], v[:, 0]))


def canonical(verts, faces):
    """Vertices sorted lexicographically; faces relabelled, each rotated so that its smallest index comes first
    (orientation preserved), rows sorted."""
    v = np.asarray(verts, np.float32)
    f = np.asarray(faces, np.int64)
    o = _order(v)
    inv = np.empty(len(v), np.int64)
    inv[o] = np.arange(len(v))
    g = inv[f]
    k = np.argmin(g, axis=1)
    r = np.stack([g[np.arange(len(g)), (k + i) % 3] for i in range(3)], 1)
    r = r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]
    return v[o], r


def topology(verts, faces):
    f = np.asarray(faces, np.int64)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    ue, counts = np.unique(e, axis=0, return_counts=True)
    par = np.arange(len(verts))

    def find(a):
        while par[a] != a:
            par[a] = par[par[a]]
            a = par[a]
        return a
    for a, b in ue:
        ra, rb = find(a), find(b)
        if ra != rb:
            par[ra] = rb
    used = np.unique(f)
    comps = len({find(a) for a in used})
    return dict(V=int(len(verts)), E=int(len(ue)), F=int(len(f)), chi=int(len(verts) - len(ue) + len(f)),
                components=comps, boundary_edges=int((counts == 1).sum()), nonmanifold_edges=int((counts > 2).sum()))


def area_volume(verts, faces):
    v = np.asarray(verts, np.float64)
    a, b, c = v[faces[:, 0]], v[faces[:, 1]], v[faces[:, 2]]
    n = np.cross(b - a, c - a)
    return 0.5 * np.linalg.norm(n, axis=1).sum(), np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0


def compare(va, fa, vb, fb, tol=1e-6):
    """Levels, strongest first: `identical` (arrays equal as returned), `triangle_set` (same canonical triangle set on
    the same vertex set: only the order / index rotation differs), `vertex_set` (same vertices up to `tol`, the
    triangulation inside polygons may differ), `topology` (V, E, F, Euler characteristic, components, boundary)."""
    out = dict(identical=bool(np.array_equal(va, vb) and np.array_equal(fa, fb)))
    ta, tb = topology(va, fa), topology(vb, fb)
    out["topology_a"], out["topology_b"] = ta, tb
    out["topology"] = all(ta[k] == tb[k] for k in ("chi", "components", "boundary_edges", "nonmanifold_edges"))
    out["same_counts"] = ta["V"] == tb["V"] and ta["F"] == tb["F"]
    out["vertex_set"] = out["triangle_set"] = False
    out["vertex_max_abs"] = None
    if len(va) == len(vb):
        ca, ra = canonical(va, fa)
        cb, rb = canonical(vb, fb)
        d = np.abs(ca.astype(np.float64) - cb.astype(np.float64)).max() if len(ca) else 0.0
        out["vertex_max_abs"] = float(d)
        out["vertex_set_exact"] = bool(np.array_equal(ca, cb))
        out["vertex_set"] = bool(d <= tol)
        out["triangle_set"] = bool(out["vertex_set"] and ra.shape == rb.shape and np.array_equal(ra, rb))
    (aa, vola), (ab, volb) = area_volume(va, fa), area_volume(vb, fb)
    out["area_rel"] = float(abs(aa - ab) / max(abs(aa), 1e-30))
    out["volume_rel"] = float(abs(vola - volb) / max(abs(vola), 1e-30))
    return out
