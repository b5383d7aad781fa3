// Marching cubes on the GPU: classify -> block prefix scan -> vertex emit -> face emit.
//
// Replaces MCSurfaceExtractor.run (Hunyuan3D-2/hy3dgen/shapegen/models/autoencoders/surface_extractors.py:67-76),
// which copies the grid to the host and runs scikit-image's single-threaded Lewiner marching cubes.
// The sequential algorithm appends vertices and faces in cell-traversal order (z, y, x with x = last array
// axis) and creates a shared edge vertex at the first cell that uses it.  Here the same order is produced
// without any sequential dependency:
//   * the first user of a grid edge is a pure function of the edge ("owner" cell, see owns_edge), so
//     vertex ids = exclusive prefix sum over cells of the number of owned vertices + rank inside the cell,
//   * face ids   = exclusive prefix sum over cells of the triangle count.
// HBM-bound integer/byte work, no tensor cores.  A warp owns 32 consecutive cells of one grid row (a "segment"); the
// grid is read ONCE, coalesced, by the classify pass: every lane loads the four points of its own x, the sign bits go
// through warp ballots, and only the lanes whose cell the surface crosses (cube index not 0 / 255) fetch the remaining
// corners and run the Lewiner tests.  That pass leaves 16 bits per cell (tiling, number of owned vertices), so the two
// emit passes never classify again: the vertex pass reads the corners of the cells that own vertices, the face pass
// reads no grid values at all.
#include <float.h>

#include "r3g_internal.h"

#define R3G_MC_TABLE_QUAL static __device__ const
#include "../../include/r3g_mc_tables.h"

namespace {

constexpr int kThreads = 256;

constexpr int kWarps = kThreads / 32;

struct McDims {
  int n0, n1, n2;   // grid points per axis (axis2 fastest)
  int c0, c1, c2;   // cells per axis
  int64_t ncells;
  unsigned segs;    // 32-cell segments per row of c2 cells
  unsigned nsegs;   // c0 * c1 * segs; segment s = (row, seg) in row-major order = the traversal order of the cells
};

struct Cell {
  int ci;       // cube index 0..255
  int til;      // tiling index
  int nt;       // triangles
  int nv;       // vertices this cell creates
};

__device__ __forceinline__ bool face_pos_connected(const double* cv, int face) {
  const unsigned char* fc = &r3g_mc_face_corner[4 * face];
  double A = cv[fc[0]], B = cv[fc[1]], C = cv[fc[2]], D = cv[fc[3]];
  // explicit roundings: no fma contraction, so the decision does not depend on the compiler's contraction choices
  double ac = __dmul_rn(A, C), bd = __dmul_rn(B, D);
  double pmn = (A > 0.0) ? __dsub_rn(ac, bd) : __dsub_rn(bd, ac);
  return pmn > -R3G_MC_EPS;
}

// Lewiner's test_interior (scikit-image: test_internal) for the sub-cases 4, 6.1, 7.4, 10.1, 12.1, 13.5: is the pair of
// same-sign corners on a body diagonal joined through the interior of the cell (then the tunnel tiling is used)?
// desc = r3g_mc_interior entry (mode, reference edge, sigma), see tools/gen_mc_tables.py.  Every product / sum is an
// explicitly rounded operation in source order (no fma contraction): the decision does not depend on the compiler.
__device__ __forceinline__ double lerp_rn(double a, double b, double t) {
  return __dadd_rn(a, __dmul_rn(__dsub_rn(b, a), t));
}
__device__ __noinline__ bool interior_joined(const double* cv, int desc) {
  const int mode = desc & 3, edge = (desc >> 2) & 15, sigma = (desc >> 6) & 1;
  double At, Bt, Ct, Dt;
  if (mode == 1) {
    const double d40 = __dsub_rn(cv[4], cv[0]), d62 = __dsub_rn(cv[6], cv[2]);
    const double d73 = __dsub_rn(cv[7], cv[3]), d51 = __dsub_rn(cv[5], cv[1]);
    const double a = __dsub_rn(__dmul_rn(d40, d62), __dmul_rn(d73, d51));
    const double b = __dsub_rn(__dsub_rn(__dadd_rn(__dmul_rn(cv[2], d40), __dmul_rn(cv[0], d62)), __dmul_rn(cv[1], d73)),
                               __dmul_rn(cv[3], d51));
    const double t = __ddiv_rn(-b, __dmul_rn(2.0, a));
    if (t < 0.0 || t > 1.0) return sigma == 0;
    At = lerp_rn(cv[0], cv[4], t);
    Bt = lerp_rn(cv[3], cv[7], t);
    Ct = lerp_rn(cv[2], cv[6], t);
    Dt = lerp_rn(cv[1], cv[5], t);
  } else {
    const int u = r3g_mc_edge_corner[2 * edge], w = r3g_mc_edge_corner[2 * edge + 1];
    const unsigned char* sl = &r3g_mc_slice[6 * edge];
    const double t = __ddiv_rn(cv[u], __dsub_rn(cv[u], cv[w]));
    At = 0.0;
    Bt = lerp_rn(cv[sl[0]], cv[sl[1]], t);
    Ct = lerp_rn(cv[sl[2]], cv[sl[3]], t);
    Dt = lerp_rn(cv[sl[4]], cv[sl[5]], t);
  }
  const int test = (At >= 0.0 ? 1 : 0) + (Bt >= 0.0 ? 2 : 0) + (Ct >= 0.0 ? 4 : 0) + (Dt >= 0.0 ? 8 : 0);
  const double acbd = __dsub_rn(__dmul_rn(At, Ct), __dmul_rn(Bt, Dt));
  bool pos_joined;
  switch (test) {
    case 7: case 11: case 13: case 14: case 15: pos_joined = true; break;
    case 5: pos_joined = !(acbd < R3G_MC_EPS); break;
    case 10: pos_joined = !(acbd >= R3G_MC_EPS); break;
    default: pos_joined = false; break;
  }
  return sigma ? pos_joined : !pos_joined;
}

// Does cell (x,y,z) create the vertex on its edge e?  (first cell in traversal order sharing the edge)
__device__ __forceinline__ bool owns_edge(int e, int x, int y, int z) {
  int info = r3g_mc_edge_info[e];
  int ox = info & 1, oy = (info >> 1) & 1, oz = (info >> 2) & 1, axis = info >> 3;
  bool okx = (axis == 0) || ox || x == 0;
  bool oky = (axis == 1) || oy || y == 0;
  bool okz = (axis == 2) || oz || z == 0;
  return okx && oky && okz;
}

__device__ __forceinline__ void load_cell(const float* __restrict__ g, const McDims& d, int x, int y, int z,
                                          float* raw) {
  const float* p = g + ((int64_t)z * d.n1 + y) * d.n2 + x;
  const int64_t sy = d.n2, sz = (int64_t)d.n1 * d.n2;
  raw[0] = __ldg(p);            raw[1] = __ldg(p + 1);
  raw[3] = __ldg(p + sy);       raw[2] = __ldg(p + sy + 1);
  raw[4] = __ldg(p + sz);       raw[5] = __ldg(p + sz + 1);
  raw[7] = __ldg(p + sz + sy);  raw[6] = __ldg(p + sz + sy + 1);
}

// (double)v - (double)level > 0  <=>  v > level for floats (the difference is exact in double), so the cube index
// is computed in float and the double-precision work below only runs for the few cells the surface crosses.
__device__ __forceinline__ int cube_index(const float* raw, float level) {
  int ci = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) ci |= (raw[i] > level) ? (1 << i) : 0;
  return ci;
}

__device__ __forceinline__ Cell eval_cell(int ci, const float* raw, float level, double* cv, int x, int y, int z) {
  Cell c;
  c.ci = ci;
  c.til = 0; c.nt = 0; c.nv = 0;
  if (c.ci == 0 || c.ci == 255) return c;
#pragma unroll
  for (int i = 0; i < 8; ++i) cv[i] = (double)raw[i] - (double)level;
  int sub = 0, j = 0;
  unsigned amb = r3g_mc_amb_faces[c.ci];
  for (int f = 0; f < 6; ++f)
    if (amb & (1u << f)) {
      if (face_pos_connected(cv, f)) sub |= 1 << j;
      ++j;
    }
  c.til = r3g_mc_tiling_offset[c.ci] + sub;
  const int idesc = r3g_mc_interior[c.til];
  if (idesc && interior_joined(cv, idesc)) c.til = r3g_mc_tunnel[c.til];
  int t0 = r3g_mc_tiling_start[c.til], t1 = r3g_mc_tiling_start[c.til + 1];
  c.nt = (t1 - t0) / 3;
  unsigned seen = 0;
  for (int t = t0; t < t1; ++t) {
    int e = r3g_mc_tri[t];
    unsigned bit = 1u << e;
    if (seen & bit) continue;
    seen |= bit;
    if (e == 12 || owns_edge(e, x, y, z)) c.nv++;
  }
  return c;
}

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(unsigned u) {
  unsigned v = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
#ifdef __CUDA_ARCH__
  return __uint_as_float(v);
#else
  float f;
  memcpy(&f, &v, 4);
  return f;
#endif
}

// Block-wide exclusive scan in thread order (one or two counters); returns the block totals.
template <int N>
__device__ __forceinline__ void block_scan(const unsigned (&v)[N], unsigned (&excl)[N], unsigned (&total)[N]) {
  __shared__ unsigned ws[N][kWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) inc[k] = v[k];
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      unsigned n = __shfl_up_sync(0xffffffffu, inc[k], o);
      if (lane >= o) inc[k] += n;
    }
  }
  if (lane == 31) {
#pragma unroll
    for (int k = 0; k < N; ++k) ws[k][w] = inc[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    unsigned off = 0, sum = 0;
#pragma unroll
    for (int i = 0; i < kWarps; ++i) {
      const unsigned t = ws[k][i];
      if (i < w) off += t;
      sum += t;
    }
    excl[k] = off + inc[k] - v[k];
    total[k] = sum;
  }
  __syncthreads();
}

__device__ __forceinline__ bool cell_coords(const McDims& d, int64_t cell, int& x, int& y, int& z) {
  if (cell >= d.ncells) return false;
  x = (int)(cell % d.c2);
  int64_t r = cell / d.c2;
  y = (int)(r % d.c1);
  z = (int)(r / d.c1);
  return true;
}

// The warp's segment: 32 consecutive cells (x0 .. x0+31) of row (y, z); warp-uniform.  Thread order inside a block =
// segment order then lane = the sequential traversal order, so block scans over threads give traversal-order ranks.
struct Segment {
  bool valid;       // the segment exists
  int x0, y, z;
};
__device__ __forceinline__ Segment warp_segment(const McDims& d) {
  Segment sg;
  const unsigned s = blockIdx.x * (unsigned)kWarps + (threadIdx.x >> 5);
  sg.valid = s < d.nsegs;
  const unsigned row = s / d.segs;
  sg.x0 = (int)(s - row * d.segs) * 32;
  sg.z = (int)(row / (unsigned)d.c1);
  sg.y = (int)(row - (unsigned)sg.z * (unsigned)d.c1);
  return sg;
}

// info word of a cell: tiling (10 bits) | number of vertices the cell creates (4 bits)
__device__ __forceinline__ unsigned short pack_info(int til, int nv) { return (unsigned short)(til | (nv << 10)); }

// Pass 1: classify.  Per block: (vertex, triangle) counts and the min / max of the points it touched (for skimage's
// level check); per cell of a block the surface crosses: the info word.
__global__ void __launch_bounds__(kThreads) mc_count_kernel(const float* __restrict__ g, McDims d, float level,
                                                            unsigned* __restrict__ block_counts,
                                                            unsigned* __restrict__ block_minmax,
                                                            unsigned short* __restrict__ info) {
  const int lane = threadIdx.x & 31;
  const Segment sg = warp_segment(d);
  const int x = sg.x0 + lane;
  const int64_t sy = d.n2, sz = (int64_t)d.n1 * d.n2;
  const float* row = g + ((int64_t)sg.z * d.n1 + sg.y) * d.n2;      // point (0, y, z)
  // own points: (x, y|y+1, z|z+1) for x <= c2 (= the last point of the row)
  float v00 = level, v10 = level, v01 = level, v11 = level, ex = level;
  float lo = INFINITY, hi = -INFINITY;
  const bool point = sg.valid && x <= d.c2;
  if (point) {
    const float* p = row + x;
    v00 = __ldg(p); v10 = __ldg(p + sy); v01 = __ldg(p + sz); v11 = __ldg(p + sz + sy);
    lo = fminf(fminf(v00, v10), fminf(v01, v11));       // fminf / fmaxf skip NaNs (FlashVDM grids carry them)
    hi = fmaxf(fmaxf(v00, v10), fmaxf(v01, v11));
  }
  // the four points at x0 + 32 (the x+1 corners of lane 31's cell): lanes 0..3 fetch one each
  const bool has_extra = sg.valid && sg.x0 + 32 <= d.c2;
  if (has_extra && lane < 4) {
    ex = __ldg(row + (lane & 1 ? sy : 0) + (lane & 2 ? sz : 0) + sg.x0 + 32);
    lo = fminf(lo, ex); hi = fmaxf(hi, ex);
  }
  const unsigned b00 = __ballot_sync(0xffffffffu, v00 > level), b10 = __ballot_sync(0xffffffffu, v10 > level);
  const unsigned b01 = __ballot_sync(0xffffffffu, v01 > level), b11 = __ballot_sync(0xffffffffu, v11 > level);
  const unsigned be = __ballot_sync(0xffffffffu, lane < 4 && ex > level);    // bit r: row r = (y + (r&1), z + (r>>1))
  // two bits per row: own point, x+1 point
  const unsigned t00 = __funnelshift_r(b00, (be >> 0) & 1u, lane) & 3u, t10 = __funnelshift_r(b10, (be >> 1) & 1u, lane) & 3u;
  const unsigned t01 = __funnelshift_r(b01, (be >> 2) & 1u, lane) & 3u, t11 = __funnelshift_r(b11, (be >> 3) & 1u, lane) & 3u;
  const int ci = (int)((t00 & 1u) | (t00 & 2u) | ((t10 & 2u) << 1) | ((t10 & 1u) << 3) | ((t01 & 1u) << 4) | ((t01 & 2u) << 4) |
                       ((t11 & 2u) << 5) | ((t11 & 1u) << 7));
  const bool cell = sg.valid && x < d.c2;
  const bool active = cell && ci != 0 && ci != 255;

  // min / max of the block (ordered-uint form, one redux per warp)
  unsigned ulo = __reduce_min_sync(0xffffffffu, f2ord(lo)), uhi = __reduce_max_sync(0xffffffffu, f2ord(hi));
  __shared__ unsigned slo[kWarps], shi[kWarps];
  if (lane == 0) { slo[threadIdx.x >> 5] = ulo; shi[threadIdx.x >> 5] = uhi; }
  const int any = __syncthreads_or(active);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < kWarps; ++i) { ulo = min(ulo, slo[i]); uhi = max(uhi, shi[i]); }
    block_minmax[2 * blockIdx.x] = ulo;       // no same-address atomics: the scan kernel reduces these
    block_minmax[2 * blockIdx.x + 1] = uhi;
  }
  if (!any) {       // the common case: the surface does not cross this block
    if (threadIdx.x == 0) { block_counts[2 * blockIdx.x] = 0; block_counts[2 * blockIdx.x + 1] = 0; }
    return;
  }
  unsigned cnt[2] = {0, 0};
  unsigned short word = 0;
  if (active) {
    const float* p = row + x;
    float raw[8];
    raw[0] = v00; raw[3] = v10; raw[4] = v01; raw[7] = v11;
    raw[1] = __ldg(p + 1); raw[2] = __ldg(p + sy + 1); raw[5] = __ldg(p + sz + 1); raw[6] = __ldg(p + sz + sy + 1);
    double cv[8];
    const Cell c = eval_cell(ci, raw, level, cv, x, sg.y, sg.z);
    cnt[0] = c.nv; cnt[1] = c.nt;
    word = pack_info(c.til, c.nv);
  }
  info[(int64_t)blockIdx.x * kThreads + threadIdx.x] = word;
  unsigned excl[2], tot[2];
  block_scan<2>(cnt, excl, tot);
  if (threadIdx.x == 0) { block_counts[2 * blockIdx.x] = tot[0]; block_counts[2 * blockIdx.x + 1] = tot[1]; }
}

// Pass 2: exclusive scan of the per-block counts and reduction of the per-block min / max.  One block of 32 warps; warp w
// owns a contiguous range of entries and walks it 32 at a time (coalesced 8-byte loads), twice: totals, then offsets.
__global__ void __launch_bounds__(1024) mc_scan_kernel(const uint2* __restrict__ counts, uint2* __restrict__ offsets,
                                                       int nblocks, const uint2* __restrict__ block_minmax,
                                                       int64_t* __restrict__ totals) {
  __shared__ unsigned long long sv[32], st[32];
  __shared__ unsigned slo[32], shi[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int per = (((nblocks + 31) / 32) + 31) & ~31;      // entries per warp, a multiple of 32
  const int begin = (int)min((long long)nblocks, (long long)w * per), end = (int)min((long long)nblocks, (long long)begin + per);
  unsigned long long av = 0, at = 0;
  unsigned lo = 0xffffffffu, hi = 0u;
  for (int i = begin + lane; i < end; i += 32) {
    const uint2 c = counts[i], m = block_minmax[i];
    av += c.x; at += c.y;
    lo = min(lo, m.x); hi = max(hi, m.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    av += __shfl_xor_sync(0xffffffffu, av, o);
    at += __shfl_xor_sync(0xffffffffu, at, o);
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if (lane == 0) { sv[w] = av; st[w] = at; slo[w] = lo; shi[w] = hi; }
  __syncthreads();
  unsigned long long run_v = 0, run_t = 0, all_v = 0, all_t = 0;
  for (int k = 0; k < 32; ++k) {
    if (k < w) { run_v += sv[k]; run_t += st[k]; }
    all_v += sv[k]; all_t += st[k];
    lo = min(lo, slo[k]); hi = max(hi, shi[k]);
  }
  for (int base = begin; base < end; base += 32) {
    const int i = base + lane;
    const uint2 c = i < end ? counts[i] : make_uint2(0u, 0u);
    unsigned iv = c.x, it = c.y;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned nv = __shfl_up_sync(0xffffffffu, iv, o), nt = __shfl_up_sync(0xffffffffu, it, o);
      if (lane >= o) { iv += nv; it += nt; }
    }
    if (i < end) offsets[i] = make_uint2((unsigned)(run_v + iv - c.x), (unsigned)(run_t + it - c.y));
    run_v += __shfl_sync(0xffffffffu, iv, 31);
    run_t += __shfl_sync(0xffffffffu, it, 31);
  }
  if (threadIdx.x == 0) {
    totals[0] = (int64_t)all_v;
    totals[1] = (int64_t)all_t;
    totals[2] = (int64_t)lo;
    totals[3] = (int64_t)hi;
  }
}

__device__ __forceinline__ int64_t edge_slot(const McDims& d, int e, int x, int y, int z) {
  if (e == 12) return 4 * (((int64_t)z * d.n1 + y) * d.n2 + x) + 3;
  int info = r3g_mc_edge_info[e];
  int gx = x + (info & 1), gy = y + ((info >> 1) & 1), gz = z + ((info >> 2) & 1);
  return 4 * (((int64_t)gz * d.n1 + gy) * d.n2 + gx) + (info >> 3);
}

struct Rescale {
  int enabled;
  double lo[3], size[3], n[3];
};

// Pass 3: every cell writes the vertices it owns, in first-use order of its tiling, and publishes their ids.
__global__ void __launch_bounds__(kThreads) mc_vertex_kernel(const float* __restrict__ g, McDims d, float level,
                                                             const unsigned* __restrict__ block_offsets,
                                                             const unsigned* __restrict__ block_counts,
                                                             const unsigned short* __restrict__ info,
                                                             int32_t* __restrict__ vid, float* __restrict__ verts,
                                                             Rescale rs) {
  if (block_counts[2 * blockIdx.x] == 0) return;  // nothing to emit in this block (the common case)
  const unsigned word = info[(int64_t)blockIdx.x * kThreads + threadIdx.x];
  const unsigned nv[1] = {word >> 10};
  unsigned excl[1], tot[1];
  block_scan<1>(nv, excl, tot);
  if (nv[0] == 0) return;
  const Segment sg = warp_segment(d);
  const int x = sg.x0 + (threadIdx.x & 31), y = sg.y, z = sg.z;
  const int til = word & 1023;
  double cv[8];
  {
    float raw[8];
    load_cell(g, d, x, y, z, raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) cv[i] = (double)raw[i] - (double)level;
  }
  unsigned id = block_offsets[2 * blockIdx.x] + excl[0];
  const int t0 = r3g_mc_tiling_start[til], t1 = r3g_mc_tiling_start[til + 1];
  unsigned seen = 0;
  for (int t = t0; t < t1; ++t) {
    const int e = r3g_mc_tri[t];
    const unsigned bit = 1u << e;
    if (seen & bit) continue;
    seen |= bit;
    if (!(e == 12 || owns_edge(e, x, y, z))) continue;
    double fx = 0, fy = 0, fz = 0, ff = 0;
    if (e == 12) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned char* o = &r3g_mc_corner_xyz[3 * i];
        double w = 1.0 / (R3G_MC_EPS + fabs(cv[i]));
        fx += o[0] * w; fy += o[1] * w; fz += o[2] * w; ff += w;
      }
    } else {
      const int a = r3g_mc_edge_corner[2 * e], b = r3g_mc_edge_corner[2 * e + 1];
      const unsigned char *oa = &r3g_mc_corner_xyz[3 * a], *ob = &r3g_mc_corner_xyz[3 * b];
      double wa = 1.0 / (R3G_MC_EPS + fabs(cv[a]));
      double wb = 1.0 / (R3G_MC_EPS + fabs(cv[b]));
      fx = oa[0] * wa + ob[0] * wb;   // offsets are 0/1: products exact, fma-safe
      fy = oa[1] * wa + ob[1] * wb;
      fz = oa[2] * wa + ob[2] * wb;
      ff = wa + wb;
    }
    float p[3];
    p[0] = (float)((double)z + fz / ff);
    p[1] = (float)((double)y + fy / ff);
    p[2] = (float)((double)x + fx / ff);
    if (rs.enabled) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        double q = __dadd_rn(__dmul_rn(__ddiv_rn((double)p[a], rs.n[a]), rs.size[a]), rs.lo[a]);
        p[a] = (float)q;
      }
    }
    verts[3 * (int64_t)id + 0] = p[0];
    verts[3 * (int64_t)id + 1] = p[1];
    verts[3 * (int64_t)id + 2] = p[2];
    vid[edge_slot(d, e, x, y, z)] = (int32_t)id;
    ++id;
  }
}

// Pass 4: faces, in cell order then tiling order, looking vertex ids up by grid edge.  Reads no grid values.
__global__ void __launch_bounds__(kThreads) mc_face_kernel(McDims d, const unsigned* __restrict__ block_offsets,
                                                           const unsigned* __restrict__ block_counts,
                                                           const unsigned short* __restrict__ info,
                                                           const int32_t* __restrict__ vid,
                                                           int32_t* __restrict__ faces) {
  if (block_counts[2 * blockIdx.x + 1] == 0) return;
  const int til = info[(int64_t)blockIdx.x * kThreads + threadIdx.x] & 1023;
  const int t0 = r3g_mc_tiling_start[til], t1 = r3g_mc_tiling_start[til + 1];
  const unsigned nt[1] = {(unsigned)(t1 - t0) / 3u};
  unsigned excl[1], tot[1];
  block_scan<1>(nt, excl, tot);
  if (nt[0] == 0) return;
  const Segment sg = warp_segment(d);
  const int x = sg.x0 + (threadIdx.x & 31);
  int64_t fo = 3 * (int64_t)(block_offsets[2 * blockIdx.x + 1] + excl[0]);
  for (int t = t0; t < t1; ++t) faces[fo++] = vid[edge_slot(d, r3g_mc_tri[t], x, sg.y, sg.z)];
}

__global__ void __launch_bounds__(kThreads) mc_case_kernel(const float* __restrict__ g, McDims d, float level,
                                                           unsigned char* __restrict__ out) {
  const int64_t cell = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int x, y, z;
  if (!cell_coords(d, cell, x, y, z)) return;
  float raw[8];
  load_cell(g, d, x, y, z, raw);
  out[cell] = r3g_mc_case[cube_index(raw, level)];
}

struct McWorkspace {
  int32_t* vid;
  unsigned short* info;
  unsigned* counts;
  unsigned* offsets;
  unsigned* minmax;
  int64_t* totals;
  int nblocks;
};

int make_dims(r3g_ctx* ctx, int n0, int n1, int n2, McDims& d) {
  if (n0 < 2 || n1 < 2 || n2 < 2) return r3g_fail(ctx, R3G_E_INVALID, "mc: grid must be at least 2x2x2");
  d.n0 = n0; d.n1 = n1; d.n2 = n2;
  d.c0 = n0 - 1; d.c1 = n1 - 1; d.c2 = n2 - 1;
  d.ncells = (int64_t)d.c0 * d.c1 * d.c2;
  d.segs = (unsigned)((d.c2 + 31) / 32);
  const int64_t nsegs = (int64_t)d.c0 * d.c1 * d.segs;
  if (nsegs > 0x7fffffffLL) return r3g_fail(ctx, R3G_E_INVALID, "mc: grid too large");
  d.nsegs = (unsigned)nsegs;
  return R3G_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int carve(r3g_ctx* ctx, const McDims& d, void* ws, size_t ws_bytes, McWorkspace& w) {
  const int64_t npts = (int64_t)d.n0 * d.n1 * d.n2;
  w.nblocks = (int)(((int64_t)d.nsegs + kWarps - 1) / kWarps);
  size_t off = 0;
  char* base = (char*)ws;
  w.vid = (int32_t*)(base + off);       off += align256(sizeof(int32_t) * 4 * (size_t)npts);
  w.info = (unsigned short*)(base + off);  off += align256(sizeof(unsigned short) * kThreads * (size_t)w.nblocks);
  w.counts = (unsigned*)(base + off);   off += align256(sizeof(unsigned) * 2 * (size_t)w.nblocks);
  w.offsets = (unsigned*)(base + off);  off += align256(sizeof(unsigned) * 2 * (size_t)w.nblocks);
  w.minmax = (unsigned*)(base + off);   off += align256(sizeof(unsigned) * 2 * (size_t)w.nblocks);
  w.totals = (int64_t*)(base + off);    off += 256;
  if (off > ws_bytes || !ws) return r3g_fail(ctx, R3G_E_WORKSPACE, "mc: workspace %zu < required %zu", ws_bytes, off);
  return R3G_OK;
}

}  // namespace

extern "C" size_t r3g_mc_workspace_bytes(int n0, int n1, int n2) {
  if (n0 < 2 || n1 < 2 || n2 < 2) return 0;
  const int64_t npts = (int64_t)n0 * n1 * n2;
  const int64_t nsegs = (int64_t)(n0 - 1) * (n1 - 1) * ((n2 - 1 + 31) / 32);
  const size_t nblocks = (size_t)((nsegs + kWarps - 1) / kWarps);
  return align256(sizeof(int32_t) * 4 * (size_t)npts) + align256(sizeof(unsigned short) * kThreads * nblocks) +
         3 * align256(sizeof(unsigned) * 2 * nblocks) + 256;
}

extern "C" int r3g_mc_count(r3g_ctx* ctx, const float* grid, int n0, int n1, int n2, float level, void* workspace,
                            size_t workspace_bytes, int64_t* nv_host, int64_t* nf_host, void* stream) {
  if (!ctx || !grid || !nv_host || !nf_host) return r3g_fail(ctx, R3G_E_INVALID, "mc_count: null argument");
  if (!ctx->encode_tiled) return r3g_fail(ctx, R3G_E_CUDA, "mc_count: no CUDA device (there is no CPU fallback)");
  r3g_device_guard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  McDims d;
  McWorkspace w;
  int rc = make_dims(ctx, n0, n1, n2, d);
  if (rc) return rc;
  rc = carve(ctx, d, workspace, workspace_bytes, w);
  if (rc) return rc;
  mc_count_kernel<<<w.nblocks, kThreads, 0, s>>>(grid, d, level, w.counts, w.minmax, w.info);
  R3G_LAUNCH_OK(ctx);
  mc_scan_kernel<<<1, 1024, 0, s>>>((const uint2*)w.counts, (uint2*)w.offsets, w.nblocks, (const uint2*)w.minmax, w.totals);
  R3G_LAUNCH_OK(ctx);
  R3G_CUDA_OK(ctx, cudaMemcpyAsync(ctx->pinned, w.totals, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  R3G_CUDA_OK(ctx, cudaStreamSynchronize(s));
  *nv_host = ctx->pinned[0];
  *nf_host = ctx->pinned[1];
  const float vmin = ord2f((unsigned)ctx->pinned[2]), vmax = ord2f((unsigned)ctx->pinned[3]);
  if (!(level >= vmin && level <= vmax)) {
    *nv_host = *nf_host = 0;
    return r3g_fail(ctx, R3G_E_LEVEL, "Surface level must be within volume data range. (level %g, range [%g, %g])",
                    level, vmin, vmax);
  }
  if (*nv_host == 0) return r3g_fail(ctx, R3G_E_NOSURFACE, "No surface found at the given iso value.");
  return R3G_OK;
}

extern "C" int r3g_mc_extract(r3g_ctx* ctx, const float* grid, int n0, int n1, int n2, float level,
                              const double* bounds_host, void* workspace, size_t workspace_bytes, float* verts,
                              int32_t* faces, void* stream) {
  if (!ctx || !grid || !verts || !faces) return r3g_fail(ctx, R3G_E_INVALID, "mc_extract: null argument");
  if (!ctx->encode_tiled) return r3g_fail(ctx, R3G_E_CUDA, "mc_extract: no CUDA device (there is no CPU fallback)");
  r3g_device_guard guard(ctx);
  cudaStream_t s = (cudaStream_t)stream;
  McDims d;
  McWorkspace w;
  int rc = make_dims(ctx, n0, n1, n2, d);
  if (rc) return rc;
  rc = carve(ctx, d, workspace, workspace_bytes, w);
  if (rc) return rc;
  Rescale rs;
  rs.enabled = bounds_host != nullptr;
  const int nax[3] = {n0, n1, n2};
  for (int a = 0; a < 3; ++a) {
    rs.lo[a] = bounds_host ? bounds_host[a] : 0.0;
    rs.size[a] = bounds_host ? bounds_host[3 + a] - bounds_host[a] : 1.0;
    rs.n[a] = (double)nax[a];
  }
  mc_vertex_kernel<<<w.nblocks, kThreads, 0, s>>>(grid, d, level, w.offsets, w.counts, w.info, w.vid, verts, rs);
  R3G_LAUNCH_OK(ctx);
  mc_face_kernel<<<w.nblocks, kThreads, 0, s>>>(d, w.offsets, w.counts, w.info, w.vid, faces);
  R3G_LAUNCH_OK(ctx);
  return R3G_OK;
}

extern "C" int r3g_mc_classify(r3g_ctx* ctx, const float* grid, int n0, int n1, int n2, float level,
                               unsigned char* case_out, void* stream) {
  if (!ctx || !grid || !case_out) return r3g_fail(ctx, R3G_E_INVALID, "mc_classify: null argument");
  if (!ctx->encode_tiled) return r3g_fail(ctx, R3G_E_CUDA, "mc_classify: no CUDA device (there is no CPU fallback)");
  r3g_device_guard guard(ctx);
  McDims d;
  int rc = make_dims(ctx, n0, n1, n2, d);
  if (rc) return rc;
  if ((d.ncells + kThreads - 1) / kThreads > 0x7fffffffLL) return r3g_fail(ctx, R3G_E_INVALID, "mc: grid too large");
  const int nblocks = (int)((d.ncells + kThreads - 1) / kThreads);
  mc_case_kernel<<<nblocks, kThreads, 0, (cudaStream_t)stream>>>(grid, d, level, case_out);
  R3G_LAUNCH_OK(ctx);
  return R3G_OK;
}
