"""ORACLE -- test infrastructure only (tests/ may import it; nothing under 3d-re-gen_b200/ may).

A SECOND, independent CPU marching cubes: every cell is evaluated directly from the rules -- trace the iso-segments on
the six faces, resolve ambiguous faces with Lewiner's test_face, chain the segments into loops, apply Lewiner's
test_interior where his switch does, triangulate -- with the rule code of tools/gen_mc_tables.py called per cell.
It does NOT read include/r3g_mc_tables.h: no table emission, no offset / sub-index arithmetic, no edge-ownership
predicate (vertices are shared through a dict keyed by grid edge, created on first use in traversal order).  So
"CUDA == oracle/mc_oracle.c" (both table driven) is no longer the only evidence: tests/test_mc_oracle.py checks
mc_oracle.c == this tracer bit for bit (vertices, faces, order) on volumes that contain every case and every sub-case.

Follows scikit-image's measure.marching_cubes(volume, level, method='lewiner') as used by
Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:69-73.  Pure Python:
use it on small volumes (<= ~33^3).
"""
import os
import sys

import numpy as np

_TOOLS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools")
if _TOOLS not in sys.path:
    sys.path.insert(0, _TOOLS)
import gen_mc_tables as G  # noqa: E402  (rule functions only; the generated header is not used)

EPS = float(np.spacing(1.0))      # scikit-image: FLT_EPSILON = np.spacing(1.0)

_CACHE = {}


def _face_pos_connected(cv, fi):
    """Lewiner test_face for the positive polarity: |AC - BD| < eps -> True, else sign(A * (AC - BD)) >= 0."""
    a, b, c, d = (cv[k] for k in G.FACE[fi])
    acbd = a * c - b * d
    if abs(acbd) < EPS:
        return True
    return a * acbd >= 0


def _interior_joined(cv, mode, edge, sigma):
    """Lewiner test_interior; True = the two same-sign corners are joined through the cell (tunnel tiling)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == 1:
            a = (cv[4] - cv[0]) * (cv[6] - cv[2]) - (cv[7] - cv[3]) * (cv[5] - cv[1])
            b = cv[2] * (cv[4] - cv[0]) + cv[0] * (cv[6] - cv[2]) - cv[1] * (cv[7] - cv[3]) - cv[3] * (cv[5] - cv[1])
            t = -b / (np.float64(2.0) * a)
            if t < 0 or t > 1:
                return sigma == 0
            at = cv[0] + (cv[4] - cv[0]) * t
            bt = cv[3] + (cv[7] - cv[3]) * t
            ct = cv[2] + (cv[6] - cv[2]) * t
            dt = cv[1] + (cv[5] - cv[1]) * t
        else:
            u, w = G.EDGE[edge]
            pu, pw = G.CORNER[u], G.CORNER[w]
            axis = [i for i in range(3) if pu[i] != pw[i]][0]
            lo, hi = [i for i in range(3) if i != axis]

            def par(axes):
                def flip(p):
                    q = list(p)
                    for a_ in axes:
                        q[a_] = 1 - q[a_]
                    return G.CORNER.index(tuple(q))
                return flip(pu), flip(pw)
            t = cv[u] / (cv[u] - cv[w])
            at = np.float64(0.0)
            (b0, b1), (c0, c1), (d0, d1) = par((lo,)), par((lo, hi)), par((hi,))
            bt = cv[b0] + (cv[b1] - cv[b0]) * t
            ct = cv[c0] + (cv[c1] - cv[c0]) * t
            dt = cv[d0] + (cv[d1] - cv[d0]) * t
    test = (1 if at >= 0 else 0) + (2 if bt >= 0 else 0) + (4 if ct >= 0 else 0) + (8 if dt >= 0 else 0)
    if test in (7, 11, 13, 14, 15):
        pos = True
    elif test == 5:
        pos = not (at * ct - bt * dt < EPS)
    elif test == 10:
        pos = not (at * ct - bt * dt >= EPS)
    else:
        pos = False
    return pos if sigma else not pos


def _cell_triangles(ci, cv):
    """Triangle list (cube-edge ids, 12 = centre vertex) of one cell, and a label of the sub-case taken."""
    amb = G.ambiguous_faces(ci)
    dec = tuple(_face_pos_connected(cv, f) for f in amb)
    key = (ci, dec)
    if key not in _CACHE:
        d = dict(zip(amb, dec))
        loops = G.loops_for(ci, d)
        it = G.interior_test(ci, d, loops)
        plain = G.triangulate(loops)
        tun = None
        if it is not None:
            _, _, _, la, lb = it
            rest = [lp for i, lp in enumerate(loops) if i not in (la, lb)]
            tun = G.tunnel(G.stored(loops[la]), G.stored(loops[lb])) + G.triangulate(rest)
        _CACHE[key] = (plain, tun, it)
    plain, tun, it = _CACHE[key]
    if it is not None and _interior_joined(cv, it[0], it[1], it[2]):
        return tun, (G.mc_case(ci), dec, "tunnel")
    return plain, (G.mc_case(ci), dec, "tested-separate" if it is not None else "plain")


def marching_cubes(volume, level=0.0, return_subcases=False):
    vol = np.ascontiguousarray(volume, dtype=np.float32)
    n0, n1, n2 = vol.shape
    if not (vol.min() <= level <= vol.max()):
        raise ValueError("Surface level must be within volume data range.")
    lvl = np.float64(np.float32(level))
    vid, verts, faces, subcases = {}, [], [], {}
    for z in range(n0 - 1):
        for y in range(n1 - 1):
            for x in range(n2 - 1):
                cv = [np.float64(vol[z + o[2], y + o[1], x + o[0]]) - lvl for o in G.CORNER]
                ci = sum((1 << i) for i in range(8) if cv[i] > 0)
                if ci in (0, 255):
                    continue
                tris, label = _cell_triangles(ci, cv)
                if return_subcases:
                    subcases[label] = subcases.get(label, 0) + 1
                for tri in tris:
                    ids = []
                    for e in tri:
                        if e == 12:
                            k = ("c", z, y, x)
                        else:
                            a, b = G.EDGE[e]
                            pa, pb = G.CORNER[a], G.CORNER[b]
                            lo = tuple(min(pa[i], pb[i]) for i in range(3))
                            axis = [i for i in range(3) if pa[i] != pb[i]][0]
                            k = (axis, z + lo[2], y + lo[1], x + lo[0])
                        if k not in vid:
                            if e == 12:
                                w = [1.0 / (EPS + abs(c)) for c in cv]
                                ff = sum(w)
                                f = [sum(G.CORNER[i][d] * w[i] for i in range(8)) for d in range(3)]
                            else:
                                wa, wb = 1.0 / (EPS + abs(cv[a])), 1.0 / (EPS + abs(cv[b]))
                                ff = wa + wb
                                f = [pa[d] * wa + pb[d] * wb for d in range(3)]
                            vid[k] = len(verts)
                            verts.append((np.float32(z + f[2] / ff), np.float32(y + f[1] / ff), np.float32(x + f[0] / ff)))
                        ids.append(vid[k])
                    faces.append(ids)
    if not verts:
        raise RuntimeError("No surface found at the given iso value.")
    v, f = np.array(verts, np.float32).reshape(-1, 3), np.array(faces, np.int32).reshape(-1, 3)
    return (v, f, subcases) if return_subcases else (v, f)
