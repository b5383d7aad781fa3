/* ORACLE -- test infrastructure only.  Nothing under 3d-re-gen_b200/ may link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * CPU restatement of the surface-extraction step of the reference:
 *   Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:67-76
 *     (MCSurfaceExtractor.run: skimage.measure.marching_cubes(vol, level, method="lewiner"),
 *      then  vertices / grid_size * bbox_size + bbox_min),
 *   surface_extractors.py:50-64 (astype(float32), ascontiguousarray(faces)).
 *
 * PARITY UNPINNED against scikit-image: the algorithm lives in scikit-image (>=0.24, requirements.txt:17),
 * which is neither vendored in /root/reference nor installed here, and the reference holds no golden
 * meshes (tests/test_mc_skimage_golden.py compares against tests/golden/skimage_mc_*.npz as soon as a machine
 * with scikit-image has run tools/dump_skimage_goldens.py).  This file follows the published algorithm and
 * scikit-image's conventions: sequential traversal `for z: for y: for x` over cells with
 * x = last array axis; cubeindex bit i set iff (v_i - level) > 0 in Lewiner's corner order; Lewiner's test_face
 * on every ambiguous face and test_interior where his switch applies it (4, 6.1, 7.4, 10.1, 12.1, 13.5);
 * one vertex per sign-changing grid edge created on first use and shared; vertex position by the
 * inverse-(eps+|value|)-weighted centre of mass in double with eps = np.spacing(1.0), stored float32; output
 * vertices in array-axis order; faces in the default gradient_direction='descent' winding.  The tiling tables are
 * the generated include/r3g_mc_tables.h (tools/gen_mc_tables.py), not Lewiner's LookUpTable.h: which sub-case a
 * cell falls into follows Lewiner's tests, the triangle order / diagonals inside a cell are our own.
 *
 * The implementation is deliberately the simple sequential one (hash-free "first user creates the
 * vertex" with per-edge id arrays), independent in structure from the scan-based CUDA kernels. */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/r3g_mc_tables.h"

typedef struct {
  float *v;   /* [nv][3] */
  int32_t *f; /* [nf][3] */
  int64_t nv, nf, capv, capf;
} mesh_t;

static int push_vertex(mesh_t *m, float a0, float a1, float a2) {
  if (m->nv == m->capv) {
    m->capv = m->capv ? m->capv * 2 : 1024;
    m->v = (float *)realloc(m->v, sizeof(float) * 3 * (size_t)m->capv);
  }
  m->v[3 * m->nv + 0] = a0;
  m->v[3 * m->nv + 1] = a1;
  m->v[3 * m->nv + 2] = a2;
  return (int)(m->nv++);
}

static void push_face(mesh_t *m, int a, int b, int c) {
  if (m->nf == m->capf) {
    m->capf = m->capf ? m->capf * 2 : 2048;
    m->f = (int32_t *)realloc(m->f, sizeof(int32_t) * 3 * (size_t)m->capf);
  }
  m->f[3 * m->nf + 0] = a;
  m->f[3 * m->nf + 1] = b;
  m->f[3 * m->nf + 2] = c;
  m->nf++;
}

/* Face test (Lewiner test_face, restated symmetrically): on an ambiguous face with diagonal
 * products P (the two positive corners) and N (the two negative corners), the positive corners
 * are joined through the face iff the bilinear saddle value is >= 0  <=>  P - N >= 0; ties within
 * R3G_MC_EPS count as joined (`if fabs(AC_BD) < FLT_EPSILON: return face >= 0`).  Doubles, like the Cython core. */
static int face_pos_connected(const double *cv, int face) {
  const unsigned char *fc = &r3g_mc_face_corner[4 * face];
  double A = cv[fc[0]], B = cv[fc[1]], C = cv[fc[2]], D = cv[fc[3]];
  double ac = A * C, bd = B * D;
  double pmn = (A > 0.0) ? (ac - bd) : (bd - ac);
  return pmn > -R3G_MC_EPS;
}

/* Lewiner's test_interior (MarchingCubes.cpp; scikit-image: test_internal), restated.  desc = r3g_mc_interior entry:
 * mode 1 (cases 4, 10): the slice z = t where At*Ct - Bt*Dt is extremal, A on v0->v4, B on v3->v7, C on v2->v6,
 * D on v1->v5; t outside [0,1] answers `s > 0`.  mode 2 (cases 6, 7, 12, 13): the slice through the iso-crossing of
 * the reference edge, At = 0.  The positive corners of the slice square are joined iff >= 3 of them are positive, or
 * the diagonal pair is positive and the bilinear saddle has their sign.  sigma = 1 (s > 0): the tunnel tiling is
 * taken iff the positive corners are joined; sigma = 0 (s < 0): iff they are not.  Returns 1 for the tunnel tiling. */
static int interior_joined(const double *cv, int desc) {
  const int mode = desc & 3, edge = (desc >> 2) & 15, sigma = (desc >> 6) & 1;
  double At, Bt, Ct, Dt;
  if (mode == 1) {
    const double a = (cv[4] - cv[0]) * (cv[6] - cv[2]) - (cv[7] - cv[3]) * (cv[5] - cv[1]);
    const double b = cv[2] * (cv[4] - cv[0]) + cv[0] * (cv[6] - cv[2]) - cv[1] * (cv[7] - cv[3]) -
                     cv[3] * (cv[5] - cv[1]);
    const double t = -b / (2.0 * a);
    if (t < 0.0 || t > 1.0) return sigma == 0;   /* `return s > 0`: "separate" for s > 0, the tunnel for s < 0 */
    At = cv[0] + (cv[4] - cv[0]) * t;
    Bt = cv[3] + (cv[7] - cv[3]) * t;
    Ct = cv[2] + (cv[6] - cv[2]) * t;
    Dt = cv[1] + (cv[5] - cv[1]) * t;
  } else {
    const int u = r3g_mc_edge_corner[2 * edge], w = r3g_mc_edge_corner[2 * edge + 1];
    const unsigned char *sl = &r3g_mc_slice[6 * edge];
    const double t = cv[u] / (cv[u] - cv[w]);
    At = 0.0;
    Bt = cv[sl[0]] + (cv[sl[1]] - cv[sl[0]]) * t;
    Ct = cv[sl[2]] + (cv[sl[3]] - cv[sl[2]]) * t;
    Dt = cv[sl[4]] + (cv[sl[5]] - cv[sl[4]]) * t;
  }
  int test = 0;
  if (At >= 0.0) test += 1;
  if (Bt >= 0.0) test += 2;
  if (Ct >= 0.0) test += 4;
  if (Dt >= 0.0) test += 8;
  int pos_joined;
  switch (test) {
    case 7: case 11: case 13: case 14: case 15: pos_joined = 1; break;
    case 5: pos_joined = !(At * Ct - Bt * Dt < R3G_MC_EPS); break;
    case 10: pos_joined = !(At * Ct - Bt * Dt >= R3G_MC_EPS); break;
    default: pos_joined = 0; break;
  }
  return sigma ? pos_joined : !pos_joined;
}

/* Returns 0 on success, 1 if level is outside [min,max] (skimage ValueError), 2 if no surface
 * (skimage RuntimeError).  vol is [n0][n1][n2] C-contiguous float32.
 * If bounds != NULL (6 doubles: min xyz, max xyz) the Hunyuan rescale
 *   v / (n_axis) * (max-min) + min   (float64 intermediate, float32 result;
 *   surface_extractors.py:74-75 divides by grid_size = R+1 = n_axis) is applied.
 * case_out (optional, [(n0-1)*(n1-1)*(n2-1)]) receives the base case 0..14 of every cell. */
int r3g_oracle_marching_cubes(const float *vol, int n0, int n1, int n2, float level, const double *bounds,
                              float **verts_out, int64_t *nv_out, int32_t **faces_out, int64_t *nf_out,
                              unsigned char *case_out) {
  const int64_t npts = (int64_t)n0 * n1 * n2;
  float vmin = INFINITY, vmax = -INFINITY;
  for (int64_t i = 0; i < npts; ++i) {
    if (vol[i] < vmin) vmin = vol[i];
    if (vol[i] > vmax) vmax = vol[i];
  }
  *verts_out = NULL;
  *faces_out = NULL;
  *nv_out = *nf_out = 0;
  if (!(level >= vmin && level <= vmax)) return 1;

  /* vertex id per grid edge: slot 0/1/2 = edge leaving the grid point along x(axis2)/y(axis1)/z(axis0),
   * slot 3 = centre vertex of the cell whose origin is the grid point. */
  int32_t *vid = (int32_t *)malloc(sizeof(int32_t) * 4 * (size_t)npts);
  memset(vid, 0xFF, sizeof(int32_t) * 4 * (size_t)npts);
  mesh_t m;
  memset(&m, 0, sizeof(m));
  const int cx = n2 - 1, cy = n1 - 1, cz = n0 - 1;

  for (int z = 0; z < cz; ++z)
    for (int y = 0; y < cy; ++y)
      for (int x = 0; x < cx; ++x) {
        double cv[8];
        int ci = 0;
        for (int i = 0; i < 8; ++i) {
          const unsigned char *o = &r3g_mc_corner_xyz[3 * i];
          float s = vol[((int64_t)(z + o[2]) * n1 + (y + o[1])) * n2 + (x + o[0])];
          cv[i] = (double)s - (double)level;
          if (cv[i] > 0.0) ci |= 1 << i;
        }
        if (case_out) case_out[((int64_t)z * cy + y) * cx + x] = r3g_mc_case[ci];
        if (ci == 0 || ci == 255) continue;
        int sub = 0, j = 0;
        unsigned amb = r3g_mc_amb_faces[ci];
        for (int f = 0; f < 6; ++f)
          if (amb & (1u << f)) {
            if (face_pos_connected(cv, f)) sub |= 1 << j;
            ++j;
          }
        int til = r3g_mc_tiling_offset[ci] + sub;
        if (r3g_mc_interior[til] && interior_joined(cv, r3g_mc_interior[til])) til = r3g_mc_tunnel[til];
        int t0 = r3g_mc_tiling_start[til], t1 = r3g_mc_tiling_start[til + 1];
        int ids[3];
        for (int t = t0; t < t1; ++t) {
          int e = r3g_mc_tri[t];
          int64_t slot;
          if (e == 12) {
            slot = 4 * (((int64_t)z * n1 + y) * n2 + x) + 3;
          } else {
            int a = r3g_mc_edge_corner[2 * e], b = r3g_mc_edge_corner[2 * e + 1];
            const unsigned char *oa = &r3g_mc_corner_xyz[3 * a], *ob = &r3g_mc_corner_xyz[3 * b];
            /* grid point = lower endpoint, axis = the coordinate that differs */
            int gx = x + (oa[0] < ob[0] ? oa[0] : ob[0]);
            int gy = y + (oa[1] < ob[1] ? oa[1] : ob[1]);
            int gz = z + (oa[2] < ob[2] ? oa[2] : ob[2]);
            int axis = (oa[0] != ob[0]) ? 0 : ((oa[1] != ob[1]) ? 1 : 2);
            slot = 4 * (((int64_t)gz * n1 + gy) * n2 + gx) + axis;
          }
          if (vid[slot] < 0) {
            double fx = 0, fy = 0, fz = 0, ff = 0;
            if (e == 12) {
              for (int i = 0; i < 8; ++i) {
                const unsigned char *o = &r3g_mc_corner_xyz[3 * i];
                double w = 1.0 / (R3G_MC_EPS + fabs(cv[i]));
                fx += o[0] * w; fy += o[1] * w; fz += o[2] * w; ff += w;
              }
            } else {
              int a = r3g_mc_edge_corner[2 * e], b = r3g_mc_edge_corner[2 * e + 1];
              const unsigned char *oa = &r3g_mc_corner_xyz[3 * a], *ob = &r3g_mc_corner_xyz[3 * b];
              double wa = 1.0 / (R3G_MC_EPS + fabs(cv[a]));
              double wb = 1.0 / (R3G_MC_EPS + fabs(cv[b]));
              fx = oa[0] * wa + ob[0] * wb;
              fy = oa[1] * wa + ob[1] * wb;
              fz = oa[2] * wa + ob[2] * wb;
              ff = wa + wb;
            }
            /* core emits (x,y,z); skimage flips to array-axis order (axis0,axis1,axis2) = (z,y,x) */
            float p0 = (float)((double)z + fz / ff);
            float p1 = (float)((double)y + fy / ff);
            float p2 = (float)((double)x + fx / ff);
            vid[slot] = push_vertex(&m, p0, p1, p2);
          }
          ids[(t - t0) % 3] = vid[slot];
          if ((t - t0) % 3 == 2) push_face(&m, ids[0], ids[1], ids[2]);
        }
      }
  free(vid);
  if (m.nv == 0) {
    free(m.v);
    free(m.f);
    return 2;
  }
  if (bounds) {
    const int nax[3] = {n0, n1, n2};
    for (int64_t i = 0; i < m.nv; ++i)
      for (int a = 0; a < 3; ++a) {
        double q = (double)m.v[3 * i + a] / (double)nax[a] * (bounds[3 + a] - bounds[a]) + bounds[a];
        m.v[3 * i + a] = (float)q;
      }
  }
  *verts_out = m.v;
  *faces_out = m.f;
  *nv_out = m.nv;
  *nf_out = m.nf;
  return 0;
}

void r3g_oracle_free(void *p) { free(p); }
