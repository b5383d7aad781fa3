#!/usr/bin/env python
"""Generate the marching-cubes tiling tables used by BOTH the CPU oracle (oracle/mc_oracle.c)
and the CUDA kernels (3d-re-gen_b200/csrc/mc.cu).

Why generated and not transcribed: the reference calls scikit-image's Lewiner marching cubes
(Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:69-73); scikit-image and
its `_marching_cubes_lewiner_luts.py` are NOT in /root/reference nor in this image, and Lewiner's
LookUpTable.h cannot be reproduced from memory.  What CAN be restated from the published
algorithm (Lewiner et al. 2003, Chernyaev 1995, Nielson-Hamann 1991) is the structure:

  * Lewiner's cube numbering (corners v0..v7, edges e0..e11, centre vertex 12, faces 1..6),
  * cubeindex = sum((v_i - level) > 0) << i,
  * ambiguous faces resolved by the bilinear-saddle face test (Lewiner `test_face`),
  * every sign-changing cube edge carries exactly one vertex, polygons of >= 8 vertices
    are tiled around the centre vertex (as Lewiner's 7.3/10.2/12.2/13.3/13.4 tilings do).

The tables here are DERIVED from those rules: for every cubeindex and every outcome of the face
tests on its ambiguous faces we trace the iso-contour segments on the six faces, chain them into
closed loops and triangulate each loop.  Interior ("tunnel") tests of MC33 (sub-cases 4.2, 6.1.2,
7.4.2, 10.1.2, 12.1.2, 13.5.2) are not generated -- see DESIGN.md "MC: what is and is not pinned".

Output: include/r3g_mc_tables.h (plain C arrays, included by the oracle and by the kernels).
"""
import itertools
import os
import sys

CORNER = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]
EDGE = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
# Lewiner's face list (test_face): corners A,B,C,D in cyclic order; A,C and B,D are the diagonals.
FACE = [(0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0), (0, 3, 2, 1), (4, 7, 6, 5)]
EDGE_OF = {}
for ei, (a, b) in enumerate(EDGE):
    EDGE_OF[(a, b)] = ei
    EDGE_OF[(b, a)] = ei


def cross(u, v):
    return (u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0])


def sub(u, v):
    return tuple(a - b for a, b in zip(u, v))


def dot(u, v):
    return sum(a * b for a, b in zip(u, v))


def ccw_from_outside(face):
    """Return the face's corners ordered counter-clockwise when seen from outside the cube."""
    p = [CORNER[c] for c in face]
    n = cross(sub(p[1], p[0]), sub(p[2], p[1]))
    centre = tuple(sum(q[i] for q in p) / 4.0 for i in range(3))
    outward = sub(centre, (0.5, 0.5, 0.5))
    return list(face) if dot(n, outward) > 0 else list(face[::-1])


FACE_CCW = [ccw_from_outside(f) for f in FACE]


def face_segments(ci, face_ccw, pos_connected):
    """Oriented iso-segments (edge_from, edge_to) on one face, positive region on the LEFT when the
    face is seen from outside the cube."""
    s = [(ci >> c) & 1 for c in face_ccw]
    npos = sum(s)
    segs = []
    if npos == 0 or npos == 4:
        return segs
    ambiguous = npos == 2 and s[0] == s[2]
    if ambiguous and pos_connected:
        # positive band through the centre: cut off each NEGATIVE corner c_k,
        # segment runs edge(c_{k-1},c_k) -> edge(c_k,c_{k+1})
        for k in range(4):
            if s[k] == 0:
                e_in = EDGE_OF[(face_ccw[(k - 1) % 4], face_ccw[k])]
                e_out = EDGE_OF[(face_ccw[k], face_ccw[(k + 1) % 4])]
                segs.append((e_in, e_out))
        return segs
    # every maximal CCW arc of positive corners c_i..c_j is cut off by one segment running
    # edge(c_j,c_{j+1}) -> edge(c_{i-1},c_i)
    for i in range(4):
        if s[i] == 1 and s[(i - 1) % 4] == 0:
            j = i
            while s[(j + 1) % 4] == 1:
                j = (j + 1) % 4
            e_from = EDGE_OF[(face_ccw[j], face_ccw[(j + 1) % 4])]
            e_to = EDGE_OF[(face_ccw[(i - 1) % 4], face_ccw[i])]
            segs.append((e_from, e_to))
    return segs


def ambiguous_faces(ci):
    out = []
    for fi, f in enumerate(FACE):
        s = [(ci >> c) & 1 for c in f]
        if sum(s) == 2 and s[0] == s[2]:
            out.append(fi)
    return out


def loops_for(ci, decisions):
    """decisions: dict face_index -> pos_connected (only for ambiguous faces)."""
    nxt = {}
    for fi in range(6):
        for (a, b) in face_segments(ci, FACE_CCW[fi], decisions.get(fi, False)):
            assert a not in nxt, "edge starts two segments"
            nxt[a] = b
    # every crossing edge must start exactly one segment and end exactly one
    crossing = [ei for ei, (a, b) in enumerate(EDGE) if ((ci >> a) & 1) != ((ci >> b) & 1)]
    assert sorted(nxt.keys()) == crossing, (ci, decisions, nxt, crossing)
    assert sorted(nxt.values()) == crossing
    loops, seen = [], set()
    for e in crossing:  # increasing edge id => loops sorted by their smallest edge, started there
        if e in seen:
            continue
        lp, cur = [], e
        while cur not in seen:
            seen.add(cur)
            lp.append(cur)
            cur = nxt[cur]
        assert cur == e
        loops.append(lp)
    return loops


def edges_share_face(e1, e2):
    a, b = set(EDGE[e1]), set(EDGE[e2])
    return any(a <= set(f) and b <= set(f) for f in FACE)


def polygon_triangulations(P):
    """All triangulations of the polygon P (vertex list, orientation preserved), fan-like ones first."""
    if len(P) < 3:
        return [[]]
    if len(P) == 3:
        return [[tuple(P)]]
    out = []
    for k in range(len(P) - 2, 0, -1):
        for left in polygon_triangulations(P[:k + 1]):
            for right in polygon_triangulations(P[k:]):
                out.append([(P[0], P[k], P[-1])] + left + right)
    return out


def triangulate(loops):
    """Triangulate every loop.  A diagonal joining two polygon vertices that lie on a common cube face
    would be generated identically by the neighbouring cell (4 triangles on one edge, flat sliver pairs), so
    only triangulations without such diagonals are admissible -- the same constraint Lewiner's hand-built
    tilings satisfy.  Polygons with no admissible triangulation, and all polygons of >= 8 vertices, are
    tiled around the centre vertex (id 12)."""
    tris = []
    for lp in loops:
        n = len(lp)
        assert n >= 3
        # The loops are traced with the positive region on the left (normal towards increasing values in the
        # core's (x,y,z) frame).  scikit-image returns vertices in array-axis order (z,y,x) -- a reflection --
        # and its default gradient_direction='descent' output has right-hand normals pointing towards
        # INCREASING values (the reason export_to_trimesh flips the faces, pipelines.py:102).  So in the
        # core frame the stored winding is the reverse of the traced one.
        lp = [lp[0]] + lp[:0:-1]
        chosen = None
        if n < 8:
            for cand in polygon_triangulations(lp):
                ok = True
                for (a, b, c) in cand:
                    for (u, v) in ((a, b), (b, c), (c, a)):
                        iu, iv = lp.index(u), lp.index(v)
                        adjacent = (iu - iv) % n in (1, n - 1)
                        if not adjacent and edges_share_face(u, v):
                            ok = False
                if ok:
                    chosen = cand
                    break
        if chosen is None:
            chosen = [(lp[i], lp[(i + 1) % n], 12) for i in range(n)]
        tris.extend(chosen)
    return tris


def mc_case(ci):
    """Lorensen/Chernyaev/Lewiner base-case number 0..14 of a cube index (shape invariant)."""
    def shape(bits):
        pts = [CORNER[i] for i in range(8) if (bits >> i) & 1]
        d = sorted(sum(abs(a - b) for a, b in zip(p, q)) for p, q in itertools.combinations(pts, 2))
        return len(pts), tuple(d)
    n, d = shape(ci)
    if n > 4:
        n, d = shape(ci ^ 0xFF)
    if n == 0:
        return 0
    if n == 1:
        return 1
    if n == 2:
        return {1: 2, 2: 3, 3: 4}[d[0]]
    if n == 3:
        return {(1, 1, 2): 5, (1, 2, 3): 6, (2, 2, 2): 7}[d]
    # n == 4
    table = {
        (1, 1, 1, 1, 2, 2): 8,    # four corners of a face
        (1, 1, 1, 2, 2, 2): 9,    # corner + its three neighbours
        (1, 1, 2, 2, 3, 3): 10,   # two opposite parallel edges
        (1, 1, 2, 2, 2, 3): 12,   # L on a face + the corner diagonal to the L's elbow
        (2, 2, 2, 2, 2, 2): 13,   # tetrahedral
        (1, 1, 1, 2, 2, 3): 11,   # zig-zag (11 and its mirror 14)
    }
    c = table[d]
    if c == 11:
        # chirality: 11 vs 14.  Path a-b-c-d of three unit steps; sign of the triple product.
        pts = [i for i in range(8) if (ci >> i) & 1]
        if len(pts) != 4:
            pts = [i for i in range(8) if not (ci >> i) & 1]
        adj = {p: [q for q in pts if q != p and sum(abs(a - b) for a, b in zip(CORNER[p], CORNER[q])) == 1]
               for p in pts}
        ends = [p for p in pts if len(adj[p]) == 1]
        a = min(ends)
        b = adj[a][0]
        cc = [q for q in adj[b] if q != a][0]
        dd = [q for q in adj[cc] if q != b][0]
        t = dot(cross(sub(CORNER[b], CORNER[a]), sub(CORNER[cc], CORNER[b])), sub(CORNER[dd], CORNER[cc]))
        # chirality convention anchored on Lewiner's table: cubeindex 23 (corners 0,1,2,4) is case 11
        c = 14 if t > 0 else 11
    return c


def main(out_path):
    amb_mask = [0] * 256
    offset = [0] * 256
    tiling_start = [0]
    tri = []
    ntil = 0
    case = [mc_case(ci) for ci in range(256)]
    for ci in range(256):
        amb = ambiguous_faces(ci)
        m = 0
        for f in amb:
            m |= 1 << f
        amb_mask[ci] = m
        offset[ci] = ntil
        for sub_id in range(1 << len(amb)):
            dec = {f: bool((sub_id >> j) & 1) for j, f in enumerate(amb)}
            t = triangulate(loops_for(ci, dec))
            for (a, b, c) in t:
                tri.extend((a, b, c))
            tiling_start.append(len(tri))
            ntil += 1
    # --- self checks -------------------------------------------------------------------------
    # (1) single positive corner v0: in the core (x,y,z) frame the normal points AWAY from v0.
    t0 = tri[tiling_start[offset[1]]:tiling_start[offset[1] + 1]]
    mid = lambda e: tuple((CORNER[EDGE[e][0]][i] + CORNER[EDGE[e][1]][i]) / 2 for i in range(3))
    a, b, c = (mid(e) for e in t0)
    nrm = cross(sub(b, a), sub(c, a))
    assert dot(nrm, sub(CORNER[0], a)) < 0, "orientation self-check failed"
    # (2) per tiling: each directed cube-surface... every polygon edge interior to the cell is used
    #     twice with opposite directions; boundary (cube-face) segments once.
    for til in range(ntil):
        tt = tri[tiling_start[til]:tiling_start[til + 1]]
        half = {}
        for k in range(0, len(tt), 3):
            v = tt[k:k + 3]
            for i in range(3):
                key = (v[i], v[(i + 1) % 3])
                assert key not in half, "duplicate directed edge in tiling"
                half[key] = 1
    max_tris = max(tiling_start[i + 1] - tiling_start[i] for i in range(ntil)) // 3
    with open(out_path, "w") as f:
        w = f.write
        w("/* GENERATED by tools/gen_mc_tables.py -- do not edit.\n")
        w(" * Marching-cubes tables restated from the published Lewiner/Chernyaev rules (face tests only);\n")
        w(" * replaces scikit-image's _marching_cubes_lewiner_luts (third-party, not in the reference tree),\n")
        w(" * call site Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:69-73. */\n")
        w("#ifndef R3G_MC_TABLES_H\n#define R3G_MC_TABLES_H\n")
        w("#ifndef R3G_MC_TABLE_QUAL\n#define R3G_MC_TABLE_QUAL static const\n#endif\n")
        w(f"#define R3G_MC_NUM_TILINGS {ntil}\n#define R3G_MC_TRI_ENTRIES {len(tri)}\n#define R3G_MC_MAX_TRIS {max_tris}\n")

        def arr(ctype, name, data, per=16):
            w(f"R3G_MC_TABLE_QUAL {ctype} {name}[{len(data)}] = {{\n")
            for i in range(0, len(data), per):
                w("  " + ",".join(str(x) for x in data[i:i + per]) + ",\n")
            w("};\n")
        arr("unsigned char", "r3g_mc_case", case)
        arr("unsigned char", "r3g_mc_amb_faces", amb_mask)
        arr("unsigned short", "r3g_mc_tiling_offset", offset)
        arr("unsigned short", "r3g_mc_tiling_start", tiling_start)
        arr("unsigned char", "r3g_mc_tri", tri, per=24)
        # geometry helpers
        arr("unsigned char", "r3g_mc_edge_corner", [c for e in EDGE for c in e], per=2)
        arr("unsigned char", "r3g_mc_corner_xyz", [c for p in CORNER for c in p], per=3)
        arr("unsigned char", "r3g_mc_face_corner", [c for fc in FACE for c in fc], per=4)
        # per edge: bit0..2 = (x,y,z) offset of the lower endpoint, bits 3..4 = axis the edge runs along
        info = []
        for (a, b) in EDGE:
            pa, pb = CORNER[a], CORNER[b]
            lo = [min(pa[i], pb[i]) for i in range(3)]
            axis = [i for i in range(3) if pa[i] != pb[i]][0]
            info.append(lo[0] | (lo[1] << 1) | (lo[2] << 2) | (axis << 3))
        arr("unsigned char", "r3g_mc_edge_info", info, per=12)
        w("#endif\n")
    print(f"tilings={ntil} tri_entries={len(tri)} max_tris={max_tris}", file=sys.stderr)


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(os.path.join(here, "..", "include", "r3g_mc_tables.h"))
