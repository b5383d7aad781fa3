// r3g_linear: Y = epilogue(X . W^T + bias) on the 5th-gen tensor cores.
//
// Persistent, warp-specialised sm_100a kernel:
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B-swizzled K-major tiles of X and W into a smem ring)
//   warp 1      MMA issuer     (one elected thread: tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16;
//                               fp32 accumulators in TMEM, double-buffered so the epilogue of tile i overlaps
//                               the main loop of tile i+1)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b: one accumulator row per thread; bias, GELU, gate*y+residual,
//                               fp16/fp32 conversion, 16-byte global stores)
// Covers every nn.Linear on the hot path: DiT qkv/proj/mlp/linear1/linear2 (hunyuan3ddit.py:196-216,259-267),
// ShapeVAE c_qkv/c_proj/c_fc (attention_blocks.py:166-182,345-363), geo-decoder query_proj/c_q/c_kv/c_proj/mlp
// (attention_blocks.py:250-261,484-494).
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <cstdio>
#include "r3g_internal.h"
#include "r3g_ptx.cuh"

namespace {

using namespace r3g;

constexpr int BM = 128;
constexpr int BK = 64;  // 64 halfs = one 128-byte swizzle row
constexpr int kNumEpilogueWarps = 8;   // two per TMEM lane quadrant, each owning half of the tile's columns
constexpr int kNumThreads = 64 + 32 * kNumEpilogueWarps;

struct LinearParams {
  int M, N, K;
  const __half* bias;
  void* y;
  int64_t ldy;
  int seg_len;            // rows per segment (== M when unsegmented)
  int64_t y_seg_stride;   // output rows between segment starts
  int tiles_per_seg;
  int act, act_col0, act_col1;
  const __half* gate;
  int64_t gate_ld;
  int gate_rows;
  const __half* residual;
  const float* residual_f32;  // fp32 residual stream (VGGT blocks): y_f32 = res_f32 + ls_gamma[n] * fp16(acc + bias)
  const float* ls_gamma;
  int out_f32;
  int tiles_m, tiles_n;
  // fused per-head q/k normalisation (0 off, 1 RMS, 2 LayerNorm) of the 64-column heads in two column ranges
  int qkn_mode, qkn_q_col0, qkn_k_col0, qkn_cols;
  float qkn_eps;
  const __half *qkn_q_w, *qkn_q_b, *qkn_k_w, *qkn_k_b;
};

template <int BN>
struct Cfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                                    kNumEpilogueWarps * 768 /*epilogue operand slices*/;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcpf(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh-GELU: 0.5 x (1 + tanh(u)) == x * sigmoid(2u), u = sqrt(2/pi) (x + 0.044715 x^3).
// 5 FMA-pipe ops + 2 MUFU per element, accurate to ~1e-6 relative (well inside the fp16 output rounding).
__device__ __forceinline__ float gelu_tanh_f(float x) {
  constexpr float a = -2.f * 0.7978845608028654f * 1.4426950408889634f;  // -2 sqrt(2/pi) log2(e)
  constexpr float b = a * 0.044715f;
  const float t = x * x;
  const float z = x * fmaf(b, t, a);          // -2u log2(e)
  return x * rcpf(1.f + ex2f(z));
}
// erf-GELU: 0.5 x (1 + erf(x / sqrt 2)), erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7).
__device__ __forceinline__ float gelu_erf_f(float x) {
  const float az = fabsf(x) * 0.7071067811865476f;
  const float t = rcpf(fmaf(0.3275911f, az, 1.f));
  float pl = fmaf(1.061405429f, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  const float e = ex2f(az * az * -1.4426950408889634f);
  const float erf_abs = fmaf(-pl, e, 1.f);
  return 0.5f * x + 0.5f * fabsf(x) * erf_abs;   // x>0: 0.5x(1+erf), x<0: 0.5x(1-erf|.|)
}

// Epilogue operands.  The epilogue warps (two per scheduler) cannot hide an L2 round trip per 32-column chunk: the
// first version loaded bias / gate / residual after tcgen05.wait::ld and 48 % of the kernel's stall samples sat on
// those loads (profiles/README.md r1f).  Now each epilogue warp stages its 128-column slice of the bias row and of
// the (at most two) gate rows its 32 rows can belong to in its own 768 bytes of shared memory while it waits for the
// accumulator, and the fp16 residual -- the only per-thread operand -- is fetched one chunk ahead of its use.
constexpr int kEpiSmemPerWarp = 768;   // uint4 [0,16) bias, [16,32) gate row of lane 0's batch, [32,48) of lane 31's
struct ChunkOperands {
  uint4 res[4];
};
__device__ __forceinline__ void prefetch_residual(const LinearParams& p, int64_t out_row, int n0, ChunkOperands& o) {
  if (p.residual) {
    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + out_row * p.ldy + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) o.res[q] = rp[q];
  }
}
// stage this warp's bias / gate slices for the tile (columns [ncol0, ncol0 + ncols)); returns this lane's gate slice
// selector (0 / 1) or -1 when gates are read from global memory (gate_rows < 64: more than two batches per warp)
__device__ __forceinline__ int stage_tile_operands(const LinearParams& p, uint4* wsm, int lane, int ncol0, int ncols,
                                                   int r_lane0, int r_lane31, int r_mine) {
  __syncwarp();   // the previous tile's readers are done
  const bool in = lane < ncols / 8 && ncol0 + lane * 8 < p.N;
  int sel = -1;
  if (p.bias && in) wsm[lane] = __ldg(reinterpret_cast<const uint4*>(p.bias + ncol0) + lane);
  if (p.gate && p.gate_rows >= 64) {
    const int b_lo = r_lane0 / p.gate_rows, b_hi = r_lane31 / p.gate_rows;
    if (in) {
      wsm[16 + lane] = __ldg(reinterpret_cast<const uint4*>(p.gate + (int64_t)b_lo * p.gate_ld + ncol0) + lane);
      wsm[32 + lane] = __ldg(reinterpret_cast<const uint4*>(p.gate + (int64_t)b_hi * p.gate_ld + ncol0) + lane);
    }
    sel = r_mine / p.gate_rows - b_lo;
  }
  __syncwarp();
  return sel;
}

// One 32-column chunk of one accumulator row: v = fp32 accumulators of columns [n0, n0+32) of logical row r.
// sbias / sgate: this chunk's 4 uint4 of the staged slices (sgate NULL: gate from global memory or no gate);
// `pre` holds this chunk's residual and is refilled for chunk next_n0 (>= 0) as soon as it has been consumed.
__device__ __forceinline__ void epilogue_chunk(const LinearParams& p, const uint32_t* v, int r, int64_t out_row,
                                               int n0, const uint4* sbias, const uint4* sgate, const __half* gate_row,
                                               ChunkOperands& pre, int next_n0) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (p.bias) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 b4 = sbias[q];
        const __half2* h = reinterpret_cast<const __half2*>(&b4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 t = __half22float2(h[j]);
          f[q * 8 + 2 * j] += t.x;
          f[q * 8 + 2 * j + 1] += t.y;
        }
      }
    }
    // act_col0 / act_col1 are multiples of 32 (checked at the API), so a 32-column chunk is activated as a whole:
    // one warp-uniform branch, then 32 independent element chains (per-element range checks serialise the MUFU chains)
    if (p.act && n0 >= p.act_col0 && n0 < p.act_col1) {
      if (p.act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_tanh_f(__half2float(__float2half_rn(f[j])));  // Linear output is fp16
      } else if (p.act == 3) {   // ReLU (the DPT head's convolutions, dpt_head.py:344-392)
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_erf_f(__half2float(__float2half_rn(f[j])));
      }
    }
    if (p.residual) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 r4 = pre.res[q];
        const __half2* h = reinterpret_cast<const __half2*>(&r4);
        uint4 g4 = make_uint4(0, 0, 0, 0);
        if (sgate) g4 = sgate[q];
        else if (gate_row) g4 = __ldg(reinterpret_cast<const uint4*>(gate_row + n0) + q);
        const __half2* gh = reinterpret_cast<const __half2*>(&g4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 rr = __half22float2(h[j]);
          float y0 = __half2float(__float2half_rn(f[q * 8 + 2 * j]));
          float y1 = __half2float(__float2half_rn(f[q * 8 + 2 * j + 1]));
          if (gate_row) {
            float2 gg = __half22float2(gh[j]);
            y0 = __half2float(__float2half_rn(gg.x * y0));
            y1 = __half2float(__float2half_rn(gg.y * y1));
          }
          f[q * 8 + 2 * j] = rr.x + y0;
          f[q * 8 + 2 * j + 1] = rr.y + y1;
        }
      }
      if (next_n0 >= 0) prefetch_residual(p, out_row, next_n0, pre);
    }
    if (p.residual_f32) {
      const float4* rp = reinterpret_cast<const float4*>(p.residual_f32 + out_row * p.ldy + n0);
      const float4* gp = reinterpret_cast<const float4*>(p.ls_gamma + n0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 r4 = rp[q];
        float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (p.ls_gamma) g4 = __ldg(gp + q);
        f[4 * q + 0] = r4.x + g4.x * __half2float(__float2half_rn(f[4 * q + 0]));
        f[4 * q + 1] = r4.y + g4.y * __half2float(__float2half_rn(f[4 * q + 1]));
        f[4 * q + 2] = r4.z + g4.z * __half2float(__float2half_rn(f[4 * q + 2]));
        f[4 * q + 3] = r4.w + g4.w * __half2float(__float2half_rn(f[4 * q + 3]));
      }
    }
    if (p.out_f32) {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + out_row * p.ldy + n0);
#pragma unroll
      for (int q = 0; q < 8; ++q) op[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    } else {
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.y) + out_row * p.ldy + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = pack_half2(f[8 * q + 0], f[8 * q + 1]);
        o.y = pack_half2(f[8 * q + 2], f[8 * q + 3]);
        o.z = pack_half2(f[8 * q + 4], f[8 * q + 5]);
        o.w = pack_half2(f[8 * q + 6], f[8 * q + 7]);
        op[q] = o;
      }
    }

}

// One 64-column head of one accumulator row with the q/k normalisation fused (qkn_mode): v0 / v1 = the two 32-column
// accumulator chunks of columns [n0, n0+64).  Same arithmetic as qk_norm_kernel (rowops.cu): the Linear output is an
// fp16 tensor, the statistics are fp32, RMS: (x * rrms).to(fp16) * scale; LayerNorm: (x - mean) * rstd * w + b.
__device__ __forceinline__ void epilogue_qknorm(const LinearParams& p, const uint32_t* v0, const uint32_t* v1,
                                                int64_t out_row, int n0, const __half* nw, const __half* nb,
                                                const uint4* sbias8) {
  float f[64];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    f[j] = __uint_as_float(v0[j]);
    f[32 + j] = __uint_as_float(v1[j]);
  }
  if (p.bias) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 b4 = sbias8[q];
      const __half2* h = reinterpret_cast<const __half2*>(&b4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 t = __half22float2(h[j]);
        f[q * 8 + 2 * j] += t.x;
        f[q * 8 + 2 * j + 1] += t.y;
      }
    }
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 64; j += 4) {
    f[j] = __half2float(__float2half_rn(f[j]));
    f[j + 1] = __half2float(__float2half_rn(f[j + 1]));
    f[j + 2] = __half2float(__float2half_rn(f[j + 2]));
    f[j + 3] = __half2float(__float2half_rn(f[j + 3]));
    if (p.qkn_mode == 1) {
      s0 = fmaf(f[j], f[j], s0); s1 = fmaf(f[j + 1], f[j + 1], s1);
      s2 = fmaf(f[j + 2], f[j + 2], s2); s3 = fmaf(f[j + 3], f[j + 3], s3);
    } else {
      s0 += f[j]; s1 += f[j + 1]; s2 += f[j + 2]; s3 += f[j + 3];
    }
  }
  if (p.qkn_mode == 1) {
    const float rrms = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.f / 64.f) + p.qkn_eps);
    const uint4* wp = reinterpret_cast<const uint4*>(nw);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 w4 = __ldg(wp + q);
      const __half2* h = reinterpret_cast<const __half2*>(&w4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h[j]);
        f[q * 8 + 2 * j] = __half2float(__float2half_rn(f[q * 8 + 2 * j] * rrms)) * t.x;
        f[q * 8 + 2 * j + 1] = __half2float(__float2half_rn(f[q * 8 + 2 * j + 1] * rrms)) * t.y;
      }
    }
  } else {
    const float mean = ((s0 + s1) + (s2 + s3)) * (1.f / 64.f);
    float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
      f[j] -= mean; f[j + 1] -= mean; f[j + 2] -= mean; f[j + 3] -= mean;
      q0 = fmaf(f[j], f[j], q0); q1 = fmaf(f[j + 1], f[j + 1], q1);
      q2 = fmaf(f[j + 2], f[j + 2], q2); q3 = fmaf(f[j + 3], f[j + 3], q3);
    }
    const float rstd = rsqrtf(((q0 + q1) + (q2 + q3)) * (1.f / 64.f) + p.qkn_eps);
    const uint4* wp = reinterpret_cast<const uint4*>(nw);
    const uint4* bp = reinterpret_cast<const uint4*>(nb);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint4 w4 = __ldg(wp + q);
      uint4 b4 = make_uint4(0, 0, 0, 0);
      if (nb) b4 = __ldg(bp + q);
      const __half2* hw = reinterpret_cast<const __half2*>(&w4);
      const __half2* hb = reinterpret_cast<const __half2*>(&b4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 tw = __half22float2(hw[j]), tb = __half22float2(hb[j]);
        f[q * 8 + 2 * j] = f[q * 8 + 2 * j] * rstd * tw.x + tb.x;
        f[q * 8 + 2 * j + 1] = f[q * 8 + 2 * j + 1] * rstd * tw.y + tb.y;
      }
    }
  }
  uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.y) + out_row * p.ldy + n0);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 o;
    o.x = pack_half2(f[8 * q + 0], f[8 * q + 1]);
    o.y = pack_half2(f[8 * q + 2], f[8 * q + 3]);
    o.z = pack_half2(f[8 * q + 4], f[8 * q + 5]);
    o.w = pack_half2(f[8 * q + 6], f[8 * q + 7]);
    op[q] = o;
  }
}

// which normalisation range (0 = q, 1 = k, -1 = none) the 64-column group starting at n0 belongs to
__device__ __forceinline__ int qkn_range(const LinearParams& p, int n0) {
  if (!p.qkn_mode) return -1;
  if (n0 >= p.qkn_q_col0 && n0 < p.qkn_q_col0 + p.qkn_cols) return 0;
  if (p.qkn_k_col0 >= 0 && n0 >= p.qkn_k_col0 && n0 < p.qkn_k_col0 + p.qkn_cols) return 1;
  return -1;
}

template <int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_w0,
              const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_w1,
              const __grid_constant__ LinearParams p0, const __grid_constant__ LinearParams p1) {
  // Two problems may share one persistent launch (r3g_linear_args.group_next: the img and txt streams of a
  // DoubleStreamBlock): tiles [0, tiles0) belong to p0, the rest to p1 (p1.tiles_m == 0: no second problem).
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + C::kStages * C::kStageBytes + 256;   // per-epilogue-warp operand slices

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles0 = p0.tiles_m * p0.tiles_n;
  const int num_tiles = tiles0 + p1.tiles_m * p1.tiles_n;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x0);
    tma_prefetch_desc(&tmap_w0);
    if (p1.tiles_m) {
      tma_prefetch_desc(&tmap_x1);
      tma_prefetch_desc(&tmap_w1);
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], kNumEpilogueWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::kTmemCols>(tmem_base_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_wait();      // everything above overlapped the previous kernel's tail; its results are visible from here on
  pdl_trigger();   // persistent grid, one CTA per SM: the next kernel's CTAs become resident as ours retire

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const bool second = tile >= tiles0;
        const LinearParams& p = second ? p1 : p0;
        const CUtensorMap* tmap_x = second ? &tmap_x1 : &tmap_x0;
        const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
        const int lt = second ? tile - tiles0 : tile;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int sg = tm / p.tiles_per_seg, l0 = (tm % p.tiles_per_seg) * BM;
        const int num_k_blocks = (p.K + BK - 1) / BK;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kStageBytesA;
          mbar_expect_tx(&full_bar[stage], C::kStageBytes);
          tma_load_3d(sa, tmap_x, &full_bar[stage], kb * BK, l0, sg, kEvictFirst);
          tma_load_2d(sb, tmap_w, &full_bar[stage], kb * BK, tn * BN, kEvictLast);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(BM, BN, false, false);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int num_k_blocks = ((tile >= tiles0 ? p1.K : p0.K) + BK - 1) / BK;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kStageBytesA;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32, 1024, 16);
            const uint64_t db = umma_desc_sw128(sb + k * 32, 1024, 16);
            umma_ss(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem slot when the MMAs retire
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int col_half = (warp - 2) >> 2;    // which half of the tile's columns this warp owns
    const int row_in_tile = quad * 32 + lane;
    uint4* wsm = reinterpret_cast<uint4*>(epi_smem + (warp - 2) * kEpiSmemPerWarp);
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const bool second = tile >= tiles0;
      const LinearParams& p = second ? p1 : p0;
      const int lt = second ? tile - tiles0 : tile;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int sg = tm / p.tiles_per_seg;
      const int l = (tm % p.tiles_per_seg) * BM + row_in_tile;  // row inside the segment
      const bool row_ok = l < p.seg_len;
      const int r = sg * p.seg_len + l;                          // logical row
      const int64_t out_row = (int64_t)sg * p.y_seg_stride + l;
      const __half* gate_row = p.gate ? p.gate + (int64_t)(row_ok ? r / p.gate_rows : 0) * p.gate_ld : nullptr;
      // stage the warp's bias / gate slices while the accumulator is still being produced
      const int l0 = l - lane, l31 = min(l0 + 31, p.seg_len - 1);
      const int ncol0 = tn * BN + col_half * (BN / 2);
      int gsel = -1;
      if (l0 < p.seg_len)
        gsel = stage_tile_operands(p, wsm, lane, ncol0, BN / 2, sg * p.seg_len + l0, sg * p.seg_len + l31,
                                   row_ok ? r : sg * p.seg_len + l31);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      ChunkOperands pre;
      bool have_pre = false;
#pragma unroll 1
      for (int c0 = col_half * (BN / 2); c0 < (col_half + 1) * (BN / 2); c0 += 32) {
        const int n0 = tn * BN + c0;
        if (n0 >= p.N) break;  // warp-uniform
        const uint4* sbias = wsm + ((c0 - col_half * (BN / 2)) >> 3);
        uint32_t v[32];
        tmem_ld32(tmem_addr(tmem_base, quad * 32, acc * BN + c0), v);
        const int nr = (BN >= 128 && (c0 & 32) == 0) ? qkn_range(p, n0) : -1;   // warp-uniform
        if (nr >= 0) {
          uint32_t v1[32];
          tmem_ld32(tmem_addr(tmem_base, quad * 32, acc * BN + c0 + 32), v1);
          tmem_ld_wait();
          if (row_ok)
            epilogue_qknorm(p, v, v1, out_row, n0, nr ? p.qkn_k_w : p.qkn_q_w, nr ? p.qkn_k_b : p.qkn_q_b, sbias);
          c0 += 32;
          have_pre = false;
          continue;
        }
        if (!have_pre && row_ok) prefetch_residual(p, out_row, n0, pre);
        const int c1 = c0 + 32;   // the next chunk of this warp's column half, if it is an ordinary one
        const bool more = c1 < (col_half + 1) * (BN / 2) && tn * BN + c1 < p.N &&
                          !(BN >= 128 && (c1 & 32) == 0 && qkn_range(p, tn * BN + c1) >= 0);
        tmem_ld_wait();
        if (row_ok)
          epilogue_chunk(p, v, r, out_row, n0, sbias, gsel >= 0 ? sbias + 16 + 16 * gsel : nullptr, gate_row, pre,
                         more ? tn * BN + c1 : -1);
        have_pre = more;
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

// One problem of a launch: tensor maps + kernel parameters.  tile_m = rows per tile (128, or 256 for the CTA pair),
// box_n = rows of W one CTA fetches per k-block, bn = output columns per tile.
struct Problem {
  LinearParams p;
  CUtensorMap tx, tw;
};

int make_problem(r3g_ctx* ctx, const r3g_linear_args* a, int tile_m, int box_n, int bn, Problem& out) {
  LinearParams& p = out.p;
  memset(&p, 0, sizeof(p));
  if (!a) return R3G_OK;   // no second problem: tiles_m = 0
  const int seg_len = a->seg_len > 0 ? a->seg_len : a->M;
  const int nseg = a->M / seg_len;
  {
    const uint64_t dims[3] = {(uint64_t)a->K, (uint64_t)seg_len, (uint64_t)nseg};
    const uint64_t strides[3] = {2, (uint64_t)a->ldx * 2, (uint64_t)(nseg > 1 ? a->x_seg_stride : seg_len) * a->ldx * 2};
    const uint32_t box[3] = {BK, BM, 1};
    int rc = r3g_make_tmap_f16(ctx, &out.tx, a->x, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    const uint64_t strides[2] = {2, (uint64_t)a->K * 2};
    const uint32_t box[2] = {BK, (uint32_t)box_n};
    int rc = r3g_make_tmap_f16(ctx, &out.tw, a->w, 2, dims, strides, box);
    if (rc) return rc;
  }
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.bias = (const __half*)a->bias;
  p.y = a->y; p.ldy = a->ldy;
  p.seg_len = seg_len;
  p.y_seg_stride = nseg > 1 ? a->y_seg_stride : seg_len;
  p.tiles_per_seg = (seg_len + tile_m - 1) / tile_m;
  p.act = a->act; p.act_col0 = a->act_col0; p.act_col1 = a->act_col1;
  p.gate = (const __half*)a->gate; p.gate_ld = a->gate_ld; p.gate_rows = a->gate_rows > 0 ? a->gate_rows : 1;
  p.residual = a->residual_f32 ? nullptr : (const __half*)a->residual;
  p.residual_f32 = a->residual_f32 ? (const float*)a->residual : nullptr;
  p.ls_gamma = (const float*)a->ls_gamma;
  p.out_f32 = a->out_f32;
  p.qkn_mode = a->qkn_mode; p.qkn_q_col0 = a->qkn_q_col0; p.qkn_k_col0 = a->qkn_k_col0; p.qkn_cols = a->qkn_cols;
  p.qkn_eps = a->qkn_eps;
  p.qkn_q_w = (const __half*)a->qkn_q_w; p.qkn_q_b = (const __half*)a->qkn_q_b;
  p.qkn_k_w = (const __half*)a->qkn_k_w; p.qkn_k_b = (const __half*)a->qkn_k_b;
  p.tiles_m = nseg * p.tiles_per_seg;
  p.tiles_n = (a->N + bn - 1) / bn;
  return R3G_OK;
}

template <int BN>
int launch_linear(r3g_ctx* ctx, const r3g_linear_args* a, const r3g_linear_args* b, cudaStream_t s) {
  using C = Cfg<BN>;
  Problem pa, pb;
  int rc = make_problem(ctx, a, BM, BN, BN, pa);
  if (rc) return rc;
  rc = make_problem(ctx, b, BM, BN, BN, pb);
  if (rc) return rc;
  if (!b) { pb.tx = pa.tx; pb.tw = pa.tw; }
  constexpr unsigned kAttrBit = BN == 256 ? R3G_ATTR_LINEAR256 : BN == 128 ? R3G_ATTR_LINEAR128 : R3G_ATTR_LINEAR64;
  if (!(ctx->attr_done & kAttrBit)) {
    R3G_CUDA_OK(ctx, cudaFuncSetAttribute(linear_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          C::kSmemBytes));
    ctx->attr_done |= kAttrBit;
  }
  const int tiles = pa.p.tiles_m * pa.p.tiles_n + pb.p.tiles_m * pb.p.tiles_n;
  const int grid = tiles < ctx->num_sms ? tiles : ctx->num_sms;
  R3G_CUDA_OK(ctx, r3g_launch_pdl(ctx, linear_kernel<BN>, dim3(grid), dim3(kNumThreads), C::kSmemBytes, s, pa.tx, pa.tw,
                                   pb.tx, pb.tw, pa.p, pb.p));
  R3G_LAUNCH_OK(ctx);
  return R3G_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA pair (cluster of 2 on one TPC) computes a 256 x 256 tile.  Each CTA stages
// ITS 128 rows of X and ITS 128 of the 256 W rows (32 KB per k-block instead of 48 KB), the leader's single thread
// issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory, and each CTA's TMEM holds its own
// 128 accumulator rows.  One third less L2 -> SMEM traffic per FLOP than the 128 x 256 single-CTA tile, which is
// what bounds that kernel (12 TB/s of L2 bandwidth at 1.0 PFLOP/s), and room for a 6-deep ring.
constexpr int kStages2 = 6;
constexpr int kStageBytes2 = BM * BK * 2 + 128 * BK * 2;       // 16 KB of X + 16 KB of W per CTA
constexpr int kSmemBytes2 = kStages2 * kStageBytes2 + 1024 + 256 + kNumEpilogueWarps * 768;
constexpr int BN2 = 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
linear_kernel_2cta(const __grid_constant__ CUtensorMap tmap_x0, const __grid_constant__ CUtensorMap tmap_w0,
                   const __grid_constant__ CUtensorMap tmap_x1, const __grid_constant__ CUtensorMap tmap_w1,
                   const __grid_constant__ LinearParams p0, const __grid_constant__ LinearParams p1) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages2 * kStageBytes2);
  uint64_t* empty_bar = full_bar + kStages2;
  uint64_t* tmem_full_bar = empty_bar + kStages2;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;     // [2]  (leader's copy is the one the MMA warp waits on)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_smem = smem + kStages2 * kStageBytes2 + 256;   // per-epilogue-warp operand slices

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();           // rank inside the CTA pair
  const bool leader = cta_rank == 0;
  const int tiles0 = p0.tiles_m * p0.tiles_n;              // tiles [0, tiles0): problem 0, the rest: problem 1
  const int num_tiles = tiles0 + p1.tiles_m * p1.tiles_n;
  const int cluster_id = blockIdx.x / 2, num_clusters = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x0);
    tma_prefetch_desc(&tmap_w0);
    if (p1.tiles_m) {
      tma_prefetch_desc(&tmap_x1);
      tma_prefetch_desc(&tmap_w1);
    }
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(&full_bar[s], 1);    // the leader's expect_tx arrival; both CTAs' TMA bytes complete on it
      mbar_init(&empty_bar[s], 1);   // the leader's multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * kNumEpilogueWarps);  // one arrival per epilogue warp of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_base_smem);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const bool second = tile >= tiles0;
        const LinearParams& p = second ? p1 : p0;
        const CUtensorMap* tmap_x = second ? &tmap_x1 : &tmap_x0;
        const CUtensorMap* tmap_w = second ? &tmap_w1 : &tmap_w0;
        const int lt = second ? tile - tiles0 : tile;
        const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
        const int sg = tm / p.tiles_per_seg, l0 = (tm % p.tiles_per_seg) * 256 + (int)cta_rank * BM;
        const int num_k_blocks = (p.K + BK - 1) / BK;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes2;
          uint8_t* sb = sa + BM * BK * 2;
          // the peer only issues its loads: its bytes are accounted on the leader's barrier (a phase cannot
          // complete before they land because the leader armed it with the bytes of BOTH CTAs)
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * kStageBytes2);
          tma_load_3d_2sm(sa, tmap_x, &full_bar[stage], kb * BK, l0, sg, kEvictFirst);
          tma_load_2d_2sm(sb, tmap_w, &full_bar[stage], kb * BK, tn * BN2 + (int)cta_rank * 128, kEvictLast);
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc = umma_idesc_f16(256, BN2, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
        const int num_k_blocks = ((tile >= tiles0 ? p1.K : p0.K) + BK - 1) / BK;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN2;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
            const uint32_t sb = sa + BM * BK * 2;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_2sm(d_tmem, umma_desc_sw128(sa + k * 32, 1024, 16), umma_desc_sw128(sb + k * 32, 1024, 16), idesc,
                          (kb | k) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], 3);
            if (kb == num_k_blocks - 1) umma_commit_2sm(&tmem_full_bar[acc], 3);
          }
          __syncwarp();
          if (++stage == kStages2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int quad = warp & 3;
    const int col_half = (warp - 2) >> 2;
    const int row_in_tile = (int)cta_rank * BM + quad * 32 + lane;
    uint4* wsm = reinterpret_cast<uint4*>(epi_smem + (warp - 2) * kEpiSmemPerWarp);
    int it = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
      const bool second = tile >= tiles0;
      const LinearParams& p = second ? p1 : p0;
      const int lt = second ? tile - tiles0 : tile;
      const int tm = lt / p.tiles_n, tn = lt % p.tiles_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int sg = tm / p.tiles_per_seg;
      const int l = (tm % p.tiles_per_seg) * 256 + row_in_tile;
      const bool row_ok = l < p.seg_len;
      const int r = sg * p.seg_len + l;                          // logical row
      const int64_t out_row = (int64_t)sg * p.y_seg_stride + l;
      const __half* gate_row = p.gate ? p.gate + (int64_t)(row_ok ? r / p.gate_rows : 0) * p.gate_ld : nullptr;
      // stage the warp's bias / gate slices while the accumulator is still being produced
      const int l0 = l - lane, l31 = min(l0 + 31, p.seg_len - 1);
      const int ncol0 = tn * BN2 + col_half * (BN2 / 2);
      int gsel = -1;
      if (l0 < p.seg_len)
        gsel = stage_tile_operands(p, wsm, lane, ncol0, BN2 / 2, sg * p.seg_len + l0, sg * p.seg_len + l31,
                                   row_ok ? r : sg * p.seg_len + l31);
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      ChunkOperands pre;
      bool have_pre = false;
#pragma unroll 1
      for (int c0 = col_half * (BN2 / 2); c0 < (col_half + 1) * (BN2 / 2); c0 += 32) {
        const int n0 = tn * BN2 + c0;
        if (n0 >= p.N) break;  // warp-uniform
        const uint4* sbias = wsm + ((c0 - col_half * (BN2 / 2)) >> 3);
        uint32_t v[32];
        tmem_ld32(tmem_addr(tmem_base, quad * 32, acc * BN2 + c0), v);
        const int nr = (BN2 >= 128 && (c0 & 32) == 0) ? qkn_range(p, n0) : -1;   // warp-uniform
        if (nr >= 0) {
          uint32_t v1[32];
          tmem_ld32(tmem_addr(tmem_base, quad * 32, acc * BN2 + c0 + 32), v1);
          tmem_ld_wait();
          if (row_ok)
            epilogue_qknorm(p, v, v1, out_row, n0, nr ? p.qkn_k_w : p.qkn_q_w, nr ? p.qkn_k_b : p.qkn_q_b, sbias);
          c0 += 32;
          have_pre = false;
          continue;
        }
        if (!have_pre && row_ok) prefetch_residual(p, out_row, n0, pre);
        const int c1 = c0 + 32;   // the next chunk of this warp's column half, if it is an ordinary one
        const bool more = c1 < (col_half + 1) * (BN2 / 2) && tn * BN2 + c1 < p.N &&
                          !(BN2 >= 128 && (c1 & 32) == 0 && qkn_range(p, tn * BN2 + c1) >= 0);
        tmem_ld_wait();
        if (row_ok)
          epilogue_chunk(p, v, r, out_row, n0, sbias, gsel >= 0 ? sbias + 16 + 16 * gsel : nullptr, gate_row, pre,
                         more ? tn * BN2 + c1 : -1);
        have_pre = more;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty_bar[acc], 0);   // the leader's MMA warp owns the wait
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

int launch_linear_2cta(r3g_ctx* ctx, const r3g_linear_args* a, const r3g_linear_args* b, cudaStream_t s) {
  Problem pa, pb;
  int rc = make_problem(ctx, a, 256, 128, BN2, pa);
  if (rc) return rc;
  rc = make_problem(ctx, b, 256, 128, BN2, pb);
  if (rc) return rc;
  if (!b) { pb.tx = pa.tx; pb.tw = pa.tw; }
  if (!(ctx->attr_done & R3G_ATTR_LINEAR_2CTA)) {
    R3G_CUDA_OK(ctx, cudaFuncSetAttribute(linear_kernel_2cta, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes2));
    ctx->attr_done |= R3G_ATTR_LINEAR_2CTA;
  }
  // persistent grids are sized to the clusters that can be CO-RESIDENT (a GPC whose SM count is not a multiple of
  // the cluster size leaves SMs out; a cluster that has to wait for a slot would run as a second wave)
  if (!ctx->max_clusters2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctx->num_sms / 2 * 2));
    cfg.blockDim = dim3(kNumThreads);
    cfg.dynamicSmemBytes = kSmemBytes2;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, linear_kernel_2cta, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = ctx->num_sms / 2;
    }
    ctx->max_clusters2 = n < ctx->num_sms / 2 ? n : ctx->num_sms / 2;
    if (getenv("R3G_DEBUG_GEMM")) fprintf(stderr, "[r3g gemm] co-resident CTA pairs: %d\n", ctx->max_clusters2);
  }
  const int tiles = pa.p.tiles_m * pa.p.tiles_n + pb.p.tiles_m * pb.p.tiles_n;
  const int clusters = tiles < ctx->max_clusters2 ? tiles : ctx->max_clusters2;
  R3G_CUDA_OK(ctx, r3g_launch_pdl(ctx, linear_kernel_2cta, dim3(2 * clusters), dim3(kNumThreads), kSmemBytes2, s, pa.tx,
                                   pa.tw, pb.tx, pb.tw, pa.p, pb.p));
  R3G_LAUNCH_OK(ctx);
  return R3G_OK;
}

int validate_linear(r3g_ctx* ctx, const r3g_linear_args* a) {
  if (!a->x || !a->w || !a->y) return r3g_fail(ctx, R3G_E_INVALID, "linear: null argument");
  if (a->N % 32 || a->K % 8 || a->ldx % 8 || a->ldy % 8)
    return r3g_fail(ctx, R3G_E_INVALID, "linear: need N %% 32 == 0, K %% 8 == 0, ldx/ldy %% 8 == 0 (N=%d K=%d)", a->N,
                    a->K);
  if (a->seg_len > 0 && a->M % a->seg_len) return r3g_fail(ctx, R3G_E_INVALID, "linear: M must be a multiple of seg_len");
  if (a->residual && (a->out_f32 != 0) != (a->residual_f32 != 0))
    return r3g_fail(ctx, R3G_E_INVALID, "linear: an fp32 output takes an fp32 residual (residual_f32=1) and vice versa");
  if (a->ls_gamma && !a->residual_f32) return r3g_fail(ctx, R3G_E_INVALID, "linear: ls_gamma needs the fp32 residual form");
  if (a->gate && (!a->residual || a->gate_ld % 8)) return r3g_fail(ctx, R3G_E_INVALID, "linear: gate needs residual");
  if (a->act && (a->act_col0 % 32 || a->act_col1 % 32))
    return r3g_fail(ctx, R3G_E_INVALID, "linear: activation column range must be aligned to 32 columns");
  if (a->qkn_mode) {
    const bool k_on = a->qkn_k_col0 >= 0;
    if (a->qkn_mode < 0 || a->qkn_mode > 2 || a->N < 128 || a->qkn_cols <= 0 || a->qkn_cols % 64 || a->qkn_q_col0 % 64 ||
        (k_on && a->qkn_k_col0 % 64) || a->qkn_q_col0 < 0 || a->qkn_q_col0 + a->qkn_cols > a->N ||
        (k_on && a->qkn_k_col0 + a->qkn_cols > a->N) || !a->qkn_q_w || (k_on && !a->qkn_k_w))
      return r3g_fail(ctx, R3G_E_INVALID, "linear: qkn ranges must be 64-column aligned inside [0, N), N >= 128, weights set");
    if (a->residual || a->out_f32)
      return r3g_fail(ctx, R3G_E_INVALID, "linear: the fused q/k normalisation takes a plain fp16 output (no residual)");
    auto overlaps = [&](int c0) { return a->act && c0 < a->act_col1 && c0 + a->qkn_cols > a->act_col0; };
    if (overlaps(a->qkn_q_col0) || (k_on && overlaps(a->qkn_k_col0)))
      return r3g_fail(ctx, R3G_E_INVALID, "linear: normalised columns must lie outside the activation range");
  }
  return R3G_OK;
}

}  // namespace

extern "C" int r3g_linear(r3g_ctx* ctx, const r3g_linear_args* a, void* stream) {
  if (!ctx || !ctx->encode_tiled) return r3g_fail(ctx, R3G_E_CUDA, "linear: no CUDA device (there is no CPU fallback)");
  if (!a) return r3g_fail(ctx, R3G_E_INVALID, "linear: null argument");
  r3g_device_guard guard(ctx);
  const r3g_linear_args* b = (const r3g_linear_args*)a->group_next;
  if (b && b->group_next) return r3g_fail(ctx, R3G_E_INVALID, "linear: at most two problems per launch");
  if (a->M <= 0) { a = b; b = nullptr; }
  if (b && b->M <= 0) b = nullptr;
  if (!a) return R3G_OK;
  int rc = validate_linear(ctx, a);
  if (rc) return rc;
  if (b && (rc = validate_linear(ctx, b))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  // Tile choice: wave efficiency (tiles / (waves * units)) times a per-tile throughput factor measured on B200
  // (CTA-pair 256x256: 1.08 for K <= 2048 else 0.93, 128x256: 1.0, 128x128: 0.7), over the tiles of BOTH problems of a
  // grouped launch.  N = 1024 GEMMs with ~6k rows, for example, fill only 1.3 waves of 128x256 tiles; the img and txt
  // streams of a DoubleStreamBlock together (6144 + 2740 rows) fill 1.95 of 2.
  const int sms = ctx->num_sms;
  auto eff = [&](int64_t tiles, int units, double factor) {
    if (tiles <= 0) return 0.0;
    const int64_t waves = (tiles + units - 1) / units;
    return factor * (double)tiles / (double)(waves * units);
  };
  if (ctx->gemm_2cta < 0) {   // R3G_GEMM_2CTA=0 disables the CTA-pair kernel
    const char* e = getenv("R3G_GEMM_2CTA");
    ctx->gemm_2cta = (e && e[0] == '0') ? 0 : 1;
  }
  auto tiles_of = [&](const r3g_linear_args* q, int tm, int tn) -> int64_t {
    if (!q) return 0;
    const int seg = q->seg_len > 0 ? q->seg_len : q->M;
    return (int64_t)(q->M / seg) * ((seg + tm - 1) / tm) * ((q->N + tn - 1) / tn);
  };
  const int n_min = b ? (a->N < b->N ? a->N : b->N) : a->N;
  const int k_max = b ? (a->K > b->K ? a->K : b->K) : a->K;
  const bool n256 = a->N % 256 == 0 && (!b || b->N % 256 == 0);
  // measured: the CTA-pair tile wins for short K (epilogue-heavy), the single-CTA tile for K >= 4096
  const double e2 = (ctx->gemm_2cta && n256) ? eff(tiles_of(a, 256, 256) + tiles_of(b, 256, 256), sms / 2,
                                                    k_max <= 2048 ? 1.08 : 0.93) : 0.0;
  const double e256 = n_min >= 256 ? eff(tiles_of(a, 128, 256) + tiles_of(b, 128, 256), sms, 1.0) : 0.0;
  const double e128 = n_min >= 128 ? eff(tiles_of(a, 128, 128) + tiles_of(b, 128, 128), sms, 0.7) : 0.0;
  if (e2 > 0.0 && e2 >= e256 && e2 >= e128) return launch_linear_2cta(ctx, a, b, s);
  if (e256 > 0.0 && e256 >= e128) return launch_linear<256>(ctx, a, b, s);
  if (n_min >= 128) return launch_linear<128>(ctx, a, b, s);
  return launch_linear<64>(ctx, a, b, s);
}
