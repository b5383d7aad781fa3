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
// HBM-bound integer/byte work: one coalesced pass over the grid per kernel; no tensor cores.
#include <float.h>

#include "r3g_internal.h"

#define R3G_MC_TABLE_QUAL static __device__ const
#include "../../include/r3g_mc_tables.h"

namespace {

constexpr int kThreads = 256;

struct McDims {
  int n0, n1, n2;   // grid points per axis (axis2 fastest)
  int c0, c1, c2;   // cells per axis
  int64_t ncells;
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

__device__ __forceinline__ Cell eval_cell(const float* raw, float level, double* cv, int x, int y, int z) {
  Cell c;
  c.ci = cube_index(raw, level);
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

// Block-wide exclusive scan of two counters in thread order; returns block totals in tot.
__device__ __forceinline__ void block_scan2(unsigned a, unsigned b, unsigned& ea, unsigned& eb, unsigned& ta,
                                            unsigned& tb) {
  __shared__ unsigned wsa[kThreads / 32], wsb[kThreads / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned ia = a, ib = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned na = __shfl_up_sync(0xffffffffu, ia, o), nb = __shfl_up_sync(0xffffffffu, ib, o);
    if (lane >= o) { ia += na; ib += nb; }
  }
  if (lane == 31) { wsa[w] = ia; wsb[w] = ib; }
  __syncthreads();
  unsigned offa = 0, offb = 0, suma = 0, sumb = 0;
#pragma unroll
  for (int i = 0; i < kThreads / 32; ++i) {
    if (i < w) { offa += wsa[i]; offb += wsb[i]; }
    suma += wsa[i]; sumb += wsb[i];
  }
  ea = offa + ia - a;
  eb = offb + ib - b;
  ta = suma; tb = sumb;
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

// Pass 1: per-block (vertex, triangle) counts and the volume's min/max (for skimage's level check).
__global__ void __launch_bounds__(kThreads) mc_count_kernel(const float* __restrict__ g, McDims d, float level,
                                                            unsigned* __restrict__ block_counts,
                                                            float* __restrict__ block_minmax) {
  const int64_t cell = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int x, y, z;
  unsigned nv = 0, nt = 0;
  float lo = INFINITY, hi = -INFINITY;
  if (cell_coords(d, cell, x, y, z)) {
    double cv[8];
    float raw[8];
    load_cell(g, d, x, y, z, raw);
#pragma unroll
    for (int i = 0; i < 8; ++i) { lo = fminf(lo, raw[i]); hi = fmaxf(hi, raw[i]); }
    Cell c = eval_cell(raw, level, cv, x, y, z);
    nv = c.nv; nt = c.nt;
  }
  unsigned ev, et, tv, tt;
  block_scan2(nv, nt, ev, et, tv, tt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float slo[kThreads / 32], shi[kThreads / 32];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 1; i < kThreads / 32; ++i) { lo = fminf(lo, slo[i]); hi = fmaxf(hi, shi[i]); }
    block_counts[2 * blockIdx.x] = tv;
    block_counts[2 * blockIdx.x + 1] = tt;
    block_minmax[2 * blockIdx.x] = lo;      // no same-address atomics: the scan kernel reduces these
    block_minmax[2 * blockIdx.x + 1] = hi;
  }
}

// Pass 2: exclusive scan of the per-block counts (single block; the array has ncells/256 entries).
__global__ void __launch_bounds__(1024) mc_scan_kernel(const unsigned* __restrict__ counts,
                                                       unsigned* __restrict__ offsets, int nblocks,
                                                       const float* __restrict__ block_minmax,
                                                       int64_t* __restrict__ totals) {
  __shared__ unsigned long long sv[32], st[32];
  __shared__ unsigned long long carry_v, carry_t;
  if (threadIdx.x == 0) { carry_v = 0; carry_t = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float lo = INFINITY, hi = -INFINITY;
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    if (i < nblocks) { lo = fminf(lo, block_minmax[2 * i]); hi = fmaxf(hi, block_minmax[2 * i + 1]); }
    unsigned long long a = (i < nblocks) ? counts[2 * i] : 0, b = (i < nblocks) ? counts[2 * i + 1] : 0;
    unsigned long long ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned long long na = __shfl_up_sync(0xffffffffu, ia, o), nb = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += na; ib += nb; }
    }
    if (lane == 31) { sv[w] = ia; st[w] = ib; }
    __syncthreads();
    unsigned long long offa = carry_v, offb = carry_t;
    for (int k = 0; k < w; ++k) { offa += sv[k]; offb += st[k]; }
    if (i < nblocks) {
      offsets[2 * i] = (unsigned)(offa + ia - a);
      offsets[2 * i + 1] = (unsigned)(offb + ib - b);
    }
    __syncthreads();
    if (threadIdx.x == 1023) { carry_v = offa + ia; carry_t = offb + ib; }
    __syncthreads();
  }
  __shared__ float slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { slo[w] = lo; shi[w] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 32; ++k) { lo = fminf(lo, slo[k]); hi = fmaxf(hi, shi[k]); }
    totals[0] = (int64_t)carry_v;
    totals[1] = (int64_t)carry_t;
    totals[2] = (int64_t)f2ord(lo);
    totals[3] = (int64_t)f2ord(hi);
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
                                                             int32_t* __restrict__ vid, float* __restrict__ verts,
                                                             Rescale rs) {
  if (block_counts[2 * blockIdx.x] == 0) return;  // nothing to emit in this block (the common case)
  const int64_t cell = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int x = 0, y = 0, z = 0;
  double cv[8];
  Cell c; c.ci = 0; c.til = 0; c.nt = 0; c.nv = 0;
  const bool valid = cell_coords(d, cell, x, y, z);
  if (valid) {
    float raw[8];
    load_cell(g, d, x, y, z, raw);
    c = eval_cell(raw, level, cv, x, y, z);
  }
  unsigned ev, et, tv, tt;
  block_scan2(c.nv, c.nt, ev, et, tv, tt);
  if (c.nv == 0) return;
  unsigned id = block_offsets[2 * blockIdx.x] + ev;
  const int t0 = r3g_mc_tiling_start[c.til], t1 = r3g_mc_tiling_start[c.til + 1];
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

// Pass 4: faces, in cell order then tiling order, looking vertex ids up by grid edge.
__global__ void __launch_bounds__(kThreads) mc_face_kernel(const float* __restrict__ g, McDims d, float level,
                                                           const unsigned* __restrict__ block_offsets,
                                                           const unsigned* __restrict__ block_counts,
                                                           const int32_t* __restrict__ vid,
                                                           int32_t* __restrict__ faces) {
  if (block_counts[2 * blockIdx.x + 1] == 0) return;
  const int64_t cell = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  int x = 0, y = 0, z = 0;
  Cell c; c.ci = 0; c.til = 0; c.nt = 0; c.nv = 0;
  if (cell_coords(d, cell, x, y, z)) {
    double cv[8];
    float raw[8];
    load_cell(g, d, x, y, z, raw);
    c = eval_cell(raw, level, cv, x, y, z);
  }
  unsigned ev, et, tv, tt;
  block_scan2(c.nv, c.nt, ev, et, tv, tt);
  if (c.nt == 0) return;
  int64_t fo = 3 * (int64_t)(block_offsets[2 * blockIdx.x + 1] + et);
  const int t0 = r3g_mc_tiling_start[c.til], t1 = r3g_mc_tiling_start[c.til + 1];
  for (int t = t0; t < t1; ++t) faces[fo++] = vid[edge_slot(d, r3g_mc_tri[t], x, y, z)];
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
  unsigned* counts;
  unsigned* offsets;
  float* minmax;
  int64_t* totals;
  int nblocks;
};

int make_dims(r3g_ctx* ctx, int n0, int n1, int n2, McDims& d) {
  if (n0 < 2 || n1 < 2 || n2 < 2) return r3g_fail(ctx, R3G_E_INVALID, "mc: grid must be at least 2x2x2");
  d.n0 = n0; d.n1 = n1; d.n2 = n2;
  d.c0 = n0 - 1; d.c1 = n1 - 1; d.c2 = n2 - 1;
  d.ncells = (int64_t)d.c0 * d.c1 * d.c2;
  if ((d.ncells + kThreads - 1) / kThreads > 0x7fffffffLL) return r3g_fail(ctx, R3G_E_INVALID, "mc: grid too large");
  return R3G_OK;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int carve(r3g_ctx* ctx, const McDims& d, void* ws, size_t ws_bytes, McWorkspace& w) {
  const int64_t npts = (int64_t)d.n0 * d.n1 * d.n2;
  w.nblocks = (int)((d.ncells + kThreads - 1) / kThreads);
  size_t off = 0;
  char* base = (char*)ws;
  w.vid = (int32_t*)(base + off);       off += align256(sizeof(int32_t) * 4 * (size_t)npts);
  w.counts = (unsigned*)(base + off);   off += align256(sizeof(unsigned) * 2 * (size_t)w.nblocks);
  w.offsets = (unsigned*)(base + off);  off += align256(sizeof(unsigned) * 2 * (size_t)w.nblocks);
  w.minmax = (float*)(base + off);      off += align256(sizeof(float) * 2 * (size_t)w.nblocks);
  w.totals = (int64_t*)(base + off);    off += 256;
  if (off > ws_bytes || !ws) return r3g_fail(ctx, R3G_E_WORKSPACE, "mc: workspace %zu < required %zu", ws_bytes, off);
  return R3G_OK;
}

}  // namespace

extern "C" size_t r3g_mc_workspace_bytes(int n0, int n1, int n2) {
  if (n0 < 2 || n1 < 2 || n2 < 2) return 0;
  const int64_t npts = (int64_t)n0 * n1 * n2;
  const int64_t ncells = (int64_t)(n0 - 1) * (n1 - 1) * (n2 - 1);
  const size_t nblocks = (size_t)((ncells + kThreads - 1) / kThreads);
  return align256(sizeof(int32_t) * 4 * (size_t)npts) + 3 * align256(sizeof(unsigned) * 2 * nblocks) + 256;
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
  mc_count_kernel<<<w.nblocks, kThreads, 0, s>>>(grid, d, level, w.counts, w.minmax);
  R3G_LAUNCH_OK(ctx);
  mc_scan_kernel<<<1, 1024, 0, s>>>(w.counts, w.offsets, w.nblocks, w.minmax, w.totals);
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
  mc_vertex_kernel<<<w.nblocks, kThreads, 0, s>>>(grid, d, level, w.offsets, w.counts, w.vid, verts, rs);
  R3G_LAUNCH_OK(ctx);
  mc_face_kernel<<<w.nblocks, kThreads, 0, s>>>(grid, d, level, w.offsets, w.counts, w.vid, faces);
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
  const int nblocks = (int)((d.ncells + kThreads - 1) / kThreads);
  mc_case_kernel<<<nblocks, kThreads, 0, (cudaStream_t)stream>>>(grid, d, level, case_out);
  R3G_LAUNCH_OK(ctx);
  return R3G_OK;
}
