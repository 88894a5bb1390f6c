// A2/A3  softmax pooling of patch tokens into prototype tokens, forward and backward.
// Reference: medmm/modeling/models/umeml_gan.py:65-80 (PathProtoGenerator) and
// medmm/modeling/ops/attention.py:345-533 (single-head cross attention, keys = values = h).
//
// With the folded query q~ = ((c Wq^T + bq)/16) Wk (exact) only h is needed per patch:
//   S_pn = q~_p . h_n ,  a = softmax_n(S) ,  pooled_p = sum_n a_pn h_n ,  lse_p = log sum_n exp S_pn
// Backward for NB in {1,2} stacked pooling blocks that share h:
//   dA_pn = dpooled_p . h_n ,  dS_pn = a_pn (dA_pn - delta_p) ,  delta_p = dpooled_p . pooled_p
//   dq~_p = sum_n dS_pn h_n
//   dh_n  = sum_blocks sum_p (a_pn dpooled_p + dS_pn q~_p) ,  dz = dh * keep_scale * [h > 0]
//
// Forward and dq~: streaming kernels over h (R,256) bf16: one TMA producer warp fills a ring of 64-row tiles
// (128-byte swizzle), eight consumer warps run the P-wide contractions with warp-level mma.sync (the outputs
// are P x 256 accumulators over the patch rows) and fp32 online-softmax statistics.  A bag is split over
// `nsplit` CTAs; partial states are merged by a log-sum-exp kernel.
// dz (the gradient that flows back into path_net) is a pair of dense GEMMs per tile and runs on tcgen05.
#include "common.cuh"
#include "launchers.h"
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int kD = 256;
constexpr int kTM = 64;                       // patch rows per tile
constexpr int kTileBytes = kTM * kD * 2;      // 32 KB
constexpr int kBoxBytes = kTM * 128;          // one [64 rows][64 cols] swizzled box
constexpr int kCW = 8;                        // consumer warps
constexpr int kThreads = (kCW + 1) * 32;

__device__ __forceinline__ uint32_t htile_off(int row, int col) {
  return (uint32_t)((col >> 6) * kBoxBytes + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}
// [rows][256] bf16 matrix with the 16-byte chunk index XOR-ed by (row & 7)
__device__ __forceinline__ uint32_t gmat_off(int row, int col) {
  return (uint32_t)(row * 512 + (((col >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}
__device__ __forceinline__ void bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// pointer arithmetic (not an integer round trip) so the compiler keeps the shared address space
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

// rows [0,nvalid) of src (fp32, row stride 256) -> bf16 swizzled rows [row0,row0+nrows) of dst; rest zero
__device__ __forceinline__ void load_gmat_rows(uint8_t* dst, int row0, int nrows, const float* src, int nvalid,
                                               int tid, int nthreads) {
  for (int i = tid; i < nrows * 64; i += nthreads) {          // 4 floats per item
    int r = i >> 6, c = (i & 63) << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nvalid) v = *reinterpret_cast<const float4*>(src + (size_t)r * kD + c);
    uint2 pk = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    *reinterpret_cast<uint2*>(dst + gmat_off(row0 + r, c)) = pk;
  }
}

// ------------------------------------------------------------------------------------------
// shared tile phases
// ------------------------------------------------------------------------------------------
// scores for one 16-row m-tile against NT n-tiles (8 stacked rows of G each) over K = 256
template <int NT>
__device__ __forceinline__ void score_tile(float (&s)[NT][4], uint32_t tile_base, uint32_t g_base, int mt,
                                           const int (&nrow0)[NT], int lane) {
#pragma unroll
  for (int j = 0; j < NT; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
  const int arow = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll 4
  for (int kk = 0; kk < 16; kk += 2) {
    uint32_t a0[4], a1[4];
    ldsm_x4(a0, tile_base + htile_off(arow, kk * 16 + (lane >> 4) * 8));
    ldsm_x4(a1, tile_base + htile_off(arow, (kk + 1) * 16 + (lane >> 4) * 8));
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      uint32_t b[4];    // b[0..1]: k-step kk, b[2..3]: k-step kk+1
      ldsm_x4(b, g_base + gmat_off(nrow0[j] + (lane & 7), kk * 16 + (lane >> 3) * 8));
      uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
      mma_bf16_16816(s[j], a0, b0);
      mma_bf16_16816(s[j], a1, b1);
    }
  }
}

// acc[MT][4][4] += E^T (MT m-tiles of 16 stacked columns starting at ecol0) . h_tile[:, fb:fb+32], K = 64 rows
template <int MT>
__device__ __forceinline__ void weighted_sum_tile(float (&acc)[MT][4][4], uint32_t tile_base, uint32_t e_base,
                                                  int e_stride, int ecol0, int fb, int lane) {
#pragma unroll
  for (int ks = 0; ks < kTM / 16; ++ks) {
    uint32_t b[2][4];
    const int krow_b = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int pr = 0; pr < 2; ++pr) ldsm_x4_t(b[pr], tile_base + htile_off(krow_b, fb + pr * 16 + (lane >> 4) * 8));
    const int krow_a = ks * 16 + (lane & 7) + (lane >> 4) * 8;
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      uint32_t a[4];
      ldsm_x4_t(a, e_base + krow_a * e_stride + (ecol0 + mi * 16 + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        uint32_t bb[2] = {b[n >> 1][(n & 1) * 2], b[n >> 1][(n & 1) * 2 + 1]};
        mma_bf16_16816(acc[mi][n], a, bb);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct PoolFwdParams {
  const int* cu;
  const float* qt;
  long long qt_stride;     // elements between bags (0: shared queries)
  float* part_acc;         // (B, nsplit, PP, 256)
  float* part_ml;          // (B, nsplit, 2, PP)
  int P, nsplit, tiles_per_split;
};

template <int PP, int STAGES>
__global__ void __launch_bounds__(kThreads, PP <= 32 ? 2 : 1)      // two CTAs per SM hide the per-tile barrier latency
pool_fwd_kernel(const __grid_constant__ CUtensorMap tm_h, const PoolFwdParams p) {
  constexpr int NT1 = PP / 16;              // score n-tiles per warp (column half)
  constexpr int MT = PP / 16;               // m-tiles of the weighted sum
  constexpr int PT_STRIDE = PP * 2 + 16;    // bytes per row of the probability tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tiles = smem;
  uint8_t* s_q = tiles + STAGES * kTileBytes;
  uint8_t* s_pt = s_q + PP * 512;
  float* s_wmax = reinterpret_cast<float*>(s_pt + kTM * PT_STRIDE);
  float* s_alpha = s_wmax + 4 * PP;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_alpha + PP);
  uint64_t* empty = full + STAGES;

  const int b = blockIdx.y, split = blockIdx.x;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int ntiles_bag = (row_end - row_begin + kTM - 1) / kTM;
  const int t0 = split * p.tiles_per_split;
  const int ntiles = max(0, min(ntiles_bag, t0 + p.tiles_per_split) - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out_ml = p.part_ml + ((size_t)b * p.nsplit + split) * 2 * PP;
  if (ntiles == 0) {                         // neutral partial state
    if (threadIdx.x < PP) { out_ml[threadIdx.x] = -INFINITY; out_ml[PP + threadIdx.x] = 0.f; }
    return;
  }
  if (warp == kCW && lane == 0) {
    tma_prefetch_desc(&tm_h);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCW); }
    mbar_fence_init();
  }
  load_gmat_rows(s_q, 0, PP, p.qt + (size_t)b * p.qt_stride, p.P, threadIdx.x, kThreads);
  __syncthreads();

  if (warp == kCW) {
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const int stage = i % STAGES;
        mbar_wait(&empty[stage], ((i / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kTileBytes);
        uint8_t* dst = tiles + (size_t)stage * kTileBytes;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx)
          tma_load_2d(dst + bx * kBoxBytes, &tm_h, &full[stage], bx * 64, row_begin + (t0 + i) * kTM);
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int mt = warp & 3, ch = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int fb = warp * 32;
  const uint32_t q_base = smem_u32(s_q), pt_base = smem_u32(s_pt);
  int nrow0[NT1];
#pragma unroll
  for (int j = 0; j < NT1; ++j) nrow0[j] = ch * (PP / 2) + j * 8;
  float m_run[NT1][2], l_part[NT1][2];
#pragma unroll
  for (int j = 0; j < NT1; ++j) { m_run[j][0] = m_run[j][1] = -INFINITY; l_part[j][0] = l_part[j][1] = 0.f; }
  float acc[MT][4][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[mi][n][0] = acc[mi][n][1] = acc[mi][n][2] = acc[mi][n][3] = 0.f;

  for (int i = 0; i < ntiles; ++i) {
    const int stage = i % STAGES;
    mbar_wait(&full[stage], (i / STAGES) & 1);
    const uint32_t tile_base = smem_u32(tiles + (size_t)stage * kTileBytes);
    float s[NT1][4];
    score_tile<NT1>(s, tile_base, q_base, mt, nrow0, lane);
    const int r_lo = row_begin + (t0 + i) * kTM + mt * 16 + g;
    const bool ok_lo = r_lo < row_end, ok_hi = (r_lo + 8) < row_end;
#pragma unroll
    for (int j = 0; j < NT1; ++j) {
      if (!ok_lo) { s[j][0] = -INFINITY; s[j][1] = -INFINITY; }
      if (!ok_hi) { s[j][2] = -INFINITY; s[j][3] = -INFINITY; }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = fmaxf(s[j][e], s[j][e + 2]);
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
        v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16));
        if (g == 0) s_wmax[mt * PP + nrow0[j] + 2 * t + e] = v;
      }
    }
    bar_sync(1, kCW * 32);
#pragma unroll
    for (int j = 0; j < NT1; ++j) {
      float pv[4];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nrow0[j] + 2 * t + e;
        float mx = fmaxf(fmaxf(s_wmax[col], s_wmax[PP + col]), fmaxf(s_wmax[2 * PP + col], s_wmax[3 * PP + col]));
        const float m_new = fmaxf(m_run[j][e], mx);     // finite: the first tile has a valid row
        const float alpha = __expf(m_run[j][e] - m_new);
        m_run[j][e] = m_new;
        if (mt == 0 && g == 0) s_alpha[col] = alpha;
        pv[e] = __expf(s[j][e] - m_new);
        pv[e + 2] = __expf(s[j][e + 2] - m_new);
        // the MMA consumes bf16 weights: accumulate the normaliser from the same rounded values
        const float r0 = __bfloat162float(__float2bfloat16_rn(pv[e]));
        const float r1 = __bfloat162float(__float2bfloat16_rn(pv[e + 2]));
        l_part[j][e] = l_part[j][e] * alpha + r0 + r1;
      }
      *reinterpret_cast<uint32_t*>(s_pt + (mt * 16 + g) * PT_STRIDE + (nrow0[j] + 2 * t) * 2) = pack_bf16x2(pv[0], pv[1]);
      *reinterpret_cast<uint32_t*>(s_pt + (mt * 16 + g + 8) * PT_STRIDE + (nrow0[j] + 2 * t) * 2) = pack_bf16x2(pv[2], pv[3]);
    }
    bar_sync(1, kCW * 32);
#pragma unroll
    for (int mi = 0; mi < MT; ++mi) {
      const float al0 = s_alpha[mi * 16 + g], al1 = s_alpha[mi * 16 + g + 8];
#pragma unroll
      for (int n = 0; n < 4; ++n) { acc[mi][n][0] *= al0; acc[mi][n][1] *= al0; acc[mi][n][2] *= al1; acc[mi][n][3] *= al1; }
    }
    weighted_sum_tile<MT>(acc, tile_base, pt_base, PT_STRIDE, 0, fb, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }

  // ---------------- write the partial state ----------------
  bar_sync(1, kCW * 32);
#pragma unroll
  for (int j = 0; j < NT1; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float v = l_part[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) {
        s_wmax[mt * PP + nrow0[j] + 2 * t + e] = v;
        if (mt == 0) s_alpha[nrow0[j] + 2 * t + e] = m_run[j][e];
      }
    }
  bar_sync(1, kCW * 32);
  if (threadIdx.x < PP) {
    const int c = threadIdx.x;
    out_ml[c] = s_alpha[c];
    out_ml[PP + c] = s_wmax[c] + s_wmax[PP + c] + s_wmax[2 * PP + c] + s_wmax[3 * PP + c];
  }
  float* out_acc = p.part_acc + ((size_t)b * p.nsplit + split) * PP * kD;
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = fb + n * 8 + 2 * t;
      *reinterpret_cast<float2*>(out_acc + (size_t)(mi * 16 + g) * kD + col) = make_float2(acc[mi][n][0], acc[mi][n][1]);
      *reinterpret_cast<float2*>(out_acc + (size_t)(mi * 16 + g + 8) * kD + col) = make_float2(acc[mi][n][2], acc[mi][n][3]);
    }
}

// merge `nsplit` partial states per (bag, prototype): grid (P, B), 256 threads = features
__global__ void pool_merge_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml,
                                  float* __restrict__ pooled, float* __restrict__ lse, int P, int PP, int nsplit) {
  const int pi = blockIdx.x, b = blockIdx.y, f = threadIdx.x;
  const float* ml = part_ml + (size_t)b * nsplit * 2 * PP;
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) {
    const float l = ml[(size_t)s * 2 * PP + PP + pi];
    if (l > 0.f) M = fmaxf(M, ml[(size_t)s * 2 * PP + pi]);
  }
  float L = 0.f, a = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float l = ml[(size_t)s * 2 * PP + PP + pi];
    if (l > 0.f) {
      const float w = __expf(ml[(size_t)s * 2 * PP + pi] - M);
      L += l * w;
      a += w * part_acc[(((size_t)b * nsplit + s) * PP + pi) * kD + f];
    }
  }
  // an empty bag has no patches: define pooled = 0, lse = -inf
  pooled[((size_t)b * P + pi) * kD + f] = L > 0.f ? a / L : 0.f;
  if (f == 0) lse[(size_t)b * P + pi] = L > 0.f ? M + logf(L) : -INFINITY;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct PoolBwdParams {
  const int* cu;
  const float* qt[2];      long long qt_stride[2];
  const float* dpool[2];   long long dpool_stride[2];
  const float* lse[2];     // (B,P)
  const float* delta[2];   // (B,P)
  float* part_dq;          // (B, nsplit, PP, 256): dq~ of block `dq_block`
  int P, nsplit, tiles_per_split, dq_block;
};

template <int PP, int NB, int STAGES>
__global__ void __launch_bounds__(kThreads, (NB == 1 && PP <= 32) ? 2 : 1)
pool_bwd_kernel(const __grid_constant__ CUtensorMap tm_h, const PoolBwdParams p) {
  constexpr int NCOL = 2 * NB * PP;           // stacked rows of G = columns of E
  constexpr int NGW = NB * PP / 16;           // prototype groups (8 wide) per warp
  constexpr int MT = PP / 16;
  constexpr int E_STRIDE = NCOL * 2 + 16;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tiles = smem;
  uint8_t* s_g = tiles + STAGES * kTileBytes;
  uint8_t* s_e = s_g + NCOL * 512;
  float* s_lse = reinterpret_cast<float*>(s_e + kTM * E_STRIDE);
  float* s_delta = s_lse + NB * PP;
  uint64_t* full = reinterpret_cast<uint64_t*>(s_delta + NB * PP);
  uint64_t* empty = full + STAGES;

  const int b = blockIdx.y, split = blockIdx.x;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int ntiles_bag = (row_end - row_begin + kTM - 1) / kTM;
  const int t0 = split * p.tiles_per_split;
  const int ntiles = max(0, min(ntiles_bag, t0 + p.tiles_per_split) - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out_dq = p.part_dq + ((size_t)b * p.nsplit + split) * PP * kD;
  if (ntiles == 0) {
    for (int i = threadIdx.x; i < PP * kD; i += kThreads) out_dq[i] = 0.f;
    return;
  }
  if (warp == kCW && lane == 0) {
    tma_prefetch_desc(&tm_h);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kCW); }
    mbar_fence_init();
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    load_gmat_rows(s_g, k * 2 * PP, PP, p.qt[k] + (size_t)b * p.qt_stride[k], p.P, threadIdx.x, kThreads);
    load_gmat_rows(s_g, k * 2 * PP + PP, PP, p.dpool[k] + (size_t)b * p.dpool_stride[k], p.P, threadIdx.x, kThreads);
    for (int i = threadIdx.x; i < PP; i += kThreads) {
      s_lse[k * PP + i] = i < p.P ? p.lse[k][(size_t)b * p.P + i] : INFINITY;     // padded prototypes: a = 0
      s_delta[k * PP + i] = i < p.P ? p.delta[k][(size_t)b * p.P + i] : 0.f;
    }
  }
  __syncthreads();

  if (warp == kCW) {
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const int stage = i % STAGES;
        mbar_wait(&empty[stage], ((i / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kTileBytes);
        uint8_t* dst = tiles + (size_t)stage * kTileBytes;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx)
          tma_load_2d(dst + bx * kBoxBytes, &tm_h, &full[stage], bx * 64, row_begin + (t0 + i) * kTM);
      }
    }
    return;
  }

  const int mt = warp & 3, ch = warp >> 2;
  const int g = lane >> 2, t = lane & 3;
  const int fb = warp * 32;
  const uint32_t g_base = smem_u32(s_g), e_base = smem_u32(s_e);
  // this warp's prototype groups: group gi -> (block, 8-wide prototype slice)
  int nrow0[2 * NGW];         // [0,NGW): S columns (q~ rows), [NGW,2NGW): dA columns (dpooled rows)
#pragma unroll
  for (int j = 0; j < NGW; ++j) {
    const int gi = ch * NGW + j;
    const int blk = gi / (PP / 8), pg = gi % (PP / 8);
    nrow0[j] = blk * 2 * PP + pg * 8;
    nrow0[NGW + j] = blk * 2 * PP + PP + pg * 8;
  }
  float dq[MT][4][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int n = 0; n < 4; ++n) dq[mi][n][0] = dq[mi][n][1] = dq[mi][n][2] = dq[mi][n][3] = 0.f;

  for (int i = 0; i < ntiles; ++i) {
    const int stage = i % STAGES;
    mbar_wait(&full[stage], (i / STAGES) & 1);
    uint8_t* tile_ptr = tiles + (size_t)stage * kTileBytes;
    const uint32_t tile_base = smem_u32(tile_ptr);
    const int tile_row0 = row_begin + (t0 + i) * kTM;
    {
      float s[2 * NGW][4];
      score_tile<2 * NGW>(s, tile_base, g_base, mt, nrow0, lane);
      const int r_lo = tile_row0 + mt * 16 + g;
      const bool ok_lo = r_lo < row_end, ok_hi = (r_lo + 8) < row_end;
      bar_sync(1, kCW * 32);                   // everyone is done reading E of the previous tile
#pragma unroll
      for (int j = 0; j < NGW; ++j) {
        const int gi = ch * NGW + j;
        const int blk = gi / (PP / 8), pg = gi % (PP / 8);
        float ds[4], av[4];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int pi = blk * PP + pg * 8 + 2 * t + e;
          const float lse = s_lse[pi], dl = s_delta[pi];
          av[e] = ok_lo ? __expf(s[j][e] - lse) : 0.f;
          av[e + 2] = ok_hi ? __expf(s[j][e + 2] - lse) : 0.f;
          ds[e] = av[e] * (s[NGW + j][e] - dl);
          ds[e + 2] = av[e + 2] * (s[NGW + j][e + 2] - dl);
        }
        uint8_t* row_lo = s_e + (mt * 16 + g) * E_STRIDE;
        uint8_t* row_hi = row_lo + 8 * E_STRIDE;
        const int c_ds = (blk * 2 * PP + pg * 8 + 2 * t) * 2, c_a = c_ds + PP * 2;
        *reinterpret_cast<uint32_t*>(row_lo + c_ds) = pack_bf16x2(ds[0], ds[1]);
        *reinterpret_cast<uint32_t*>(row_hi + c_ds) = pack_bf16x2(ds[2], ds[3]);
        *reinterpret_cast<uint32_t*>(row_lo + c_a) = pack_bf16x2(av[0], av[1]);
        *reinterpret_cast<uint32_t*>(row_hi + c_a) = pack_bf16x2(av[2], av[3]);
      }
    }
    bar_sync(1, kCW * 32);
    // dq~ of the requested block: dS^T . h
    weighted_sum_tile<MT>(dq, tile_base, e_base, E_STRIDE, p.dq_block * 2 * PP, fb, lane);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
  }

#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int col = fb + n * 8 + 2 * t;
      *reinterpret_cast<float2*>(out_dq + (size_t)(mi * 16 + g) * kD + col) = make_float2(dq[mi][n][0], dq[mi][n][1]);
      *reinterpret_cast<float2*>(out_dq + (size_t)(mi * 16 + g + 8) * kD + col) = make_float2(dq[mi][n][2], dq[mi][n][3]);
    }
}

// out[b][pi][f] = sum_s part[b][s][pi][f]   (pi < P <= PP); grid (P, B), 256 threads
__global__ void reduce_dq_kernel(const float* __restrict__ part, float* __restrict__ out, int P, int PP, int nsplit) {
  const int pi = blockIdx.x, b = blockIdx.y, f = threadIdx.x;
  float a = 0.f;
  for (int s = 0; s < nsplit; ++s) a += part[(((size_t)b * nsplit + s) * PP + pi) * kD + f];
  out[((size_t)b * P + pi) * kD + f] = a;
}
// db1[f] (+)= sum_i part[i][f]; one block of 256 threads
__global__ void reduce_db_kernel(const float* __restrict__ part, float* __restrict__ db, int n, int accumulate) {
  const int f = threadIdx.x;
  float a = 0.f;
  for (int i = 0; i < n; ++i) a += part[(size_t)i * kD + f];
  db[f] = accumulate ? db[f] + a : a;
}

// ------------------------------------------------------------------------------------------
// dz pass on tcgen05:  dz = keep_scale * [h > 0] * (E . G),  E = [dS | a] per block from S = h . G^T
//   G (NCOL x 256, bf16, 128-byte swizzled boxes of 64 features) stacks [q~_k ; dpooled_k] of the NB blocks and
//   serves both MMAs from the same shared-memory image: K-major B operand of S = h G^T (M 128, N NCOL, K 256)
//   and MN-major B operand of dh = E G (M 128, N 256, K NCOL).  Per 128-row tile of h:
//     TMA h tile -> MMA1 -> epilogue 1 (TMEM S -> a = exp(S - lse), dS = a (dA - delta) -> E in smem, bf16)
//                -> MMA2 -> epilogue 2 (TMEM dh -> ReLU/dropout mask from the h tile, in place) -> coalesced
//     store of dz and the column sums for db1.  MMA1 of the next tile overlaps epilogue 2 of this one.
//   Warps: 0-7 epilogue (thread = patch row, warp pair splits the columns), 8 TMA producer, 9 MMA issuer.
// ------------------------------------------------------------------------------------------
constexpr int kZM = 128;                         // patch rows per tile
constexpr int kZTile = kZM * kD * 2;             // 64 KB: four [128][64] boxes
constexpr int kZThreads = 10 * 32;
__device__ __forceinline__ uint32_t box128_off(int row, int col) {   // [128 rows][64 cols] bf16 boxes, 128-B swizzle
  return (uint32_t)((col >> 6) * (kZM * 128) + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + ((col & 7) << 1));
}

struct DzParams {
  const int* cu;
  const float* qt[2];      long long qt_stride[2];
  const float* dpool[2];   long long dpool_stride[2];
  const float* lse[2];     // (B,P)
  const float* delta[2];   // (B,P)
  float* part_db;          // (B*nsplit, 256) or null
  float* part_dq;          // (B, nsplit, PP, 256) partial dq~ of block `dq_block`, or null
  bf16* dz;                // (R,256)
  int P, nsplit, tiles_per_split, dq_block;
  int relu_mask;
  float keep_scale;
};

template <int PP, int NB>
constexpr size_t dz_smem() { return 1024 + (size_t)2 * NB * PP * 512 + (size_t)kZM * ((2 * NB * PP + 63) / 64) * 128 + 2 * kZTile + 2 * NB * PP * 4 + 256; }

template <int PP, int NB>
__global__ void __launch_bounds__(kZThreads, 1)
pool_bwd_dz_kernel(const __grid_constant__ CUtensorMap tm_h, const DzParams p) {
  constexpr int NCOL = 2 * NB * PP;              // rows of G = columns of S and E
  constexpr int EBOX = (NCOL + 63) / 64;
  constexpr int NS = NB * PP / 2;                // prototype slots per epilogue thread (warp pair splits them)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_g = smem;                                        // G: 4 boxes [NCOL][64], box stride NCOL*128
  uint8_t* s_e = s_g + NCOL * 512;                            // E: EBOX boxes [128][64]
  uint8_t* tiles = s_e + EBOX * kZM * 128;                    // 2 x h tile
  float* s_lse = reinterpret_cast<float*>(tiles + 2 * kZTile);
  float* s_delta = s_lse + NB * PP;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_delta + NB * PP);
  uint64_t* full = bars;          // [2] TMA -> MMA1
  uint64_t* empty = bars + 2;     // [2] 8 epilogue warps -> TMA
  uint64_t* sfull = bars + 4;     // MMA1 -> epilogue 1
  uint64_t* sempty = bars + 5;    // 8 warps -> MMA1 (S consumed)
  uint64_t* efull = bars + 6;     // 8 warps -> MMA2 (E written)
  uint64_t* dfull = bars + 7;     // MMA2 -> epilogue 2
  uint64_t* dempty = bars + 8;    // 8 warps -> MMA2 (dh consumed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int b = blockIdx.y, split = blockIdx.x;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int ntiles_bag = (row_end - row_begin + kZM - 1) / kZM;
  const int t0 = split * p.tiles_per_split;
  const int ntiles = max(0, min(ntiles_bag, t0 + p.tiles_per_split) - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (ntiles == 0) {
    if (p.part_db)
      for (int i = threadIdx.x; i < kD; i += kZThreads) p.part_db[((size_t)b * p.nsplit + split) * kD + i] = 0.f;
    if (p.part_dq)
      for (int i = threadIdx.x; i < PP * kD; i += kZThreads) p.part_dq[((size_t)b * p.nsplit + split) * PP * kD + i] = 0.f;
    return;
  }
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_h);
    for (int i = 0; i < 2; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    mbar_init(sfull, 1); mbar_init(sempty, 8); mbar_init(efull, 8); mbar_init(dfull, 1); mbar_init(dempty, 8);
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();                               // barriers initialised, TMEM allocated: the TMA warp starts streaming h now
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp != 8) {
    // G = [q~_0 ; dpooled_0 ; q~_1 ; dpooled_1] as bf16 boxes (padded prototypes are zero rows), staged by the other
    // nine warps with the loads issued four deep (a load -> pack -> store chain per item was 10 % of the kernel)
    const int tid = warp == 9 ? 256 + lane : threadIdx.x;       // 288 workers
    constexpr int kWorkers = 288, kItems = NCOL * 64;
    for (int base = tid; base < kItems; base += 4 * kWorkers) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * kWorkers;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < kItems) {
          const int r = i >> 6, c = (i & 63) << 2;
          const int blk = r / (2 * PP), within = r % (2 * PP), pi = within % PP;
          if (pi < p.P) {
            const float* src = within >= PP ? p.dpool[blk] + (size_t)b * p.dpool_stride[blk] : p.qt[blk] + (size_t)b * p.qt_stride[blk];
            v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)pi * kD + c));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * kWorkers;
        if (i < kItems) {
          const int r = i >> 6, c = (i & 63) << 2;
          const uint32_t off = (uint32_t)((c >> 6) * (NCOL * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
          *reinterpret_cast<uint2*>(s_g + off) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
        }
      }
    }
    for (int i = tid; i < NB * PP; i += kWorkers) {
      const int blk = i / PP, pi = i % PP;
      s_lse[i] = pi < p.P ? p.lse[blk][(size_t)b * p.P + pi] : INFINITY;      // padded prototypes: a = 0
      s_delta[i] = pi < p.P ? p.delta[blk][(size_t)b * p.P + pi] : 0.f;
    }
    fence_proxy_async_smem();                    // G was written by the generic proxy, the MMAs read it through the async proxy
    bar_sync(3, kWorkers);
  }
  const uint32_t tm_s = tmem_base, tm_d = tmem_base + 128, tm_q = tmem_base + 384;   // S | dh | dq~^T (2 x 64 columns)
  // dq~ of block p.dq_block rides along: dq~^T[f][p] += sum_n h[n][f] dS[n][p] with A = h^T (the tile read MN-major) and
  // B = the 64-column box of E that holds dS of that block (the same image MMA2 reads K-major, here MN-major)
  const int dq_col = p.dq_block * 2 * PP;

  if (warp == 8) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const int stage = i & 1;
        mbar_wait_idle(&empty[stage], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kZTile);
        uint8_t* dst = tiles + (size_t)stage * kZTile;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx)
          tma_load_2d(dst + bx * (kZM * 128), &tm_h, &full[stage], bx * 64, row_begin + (t0 + i) * kZM);
      }
    }
  } else if (warp == 9) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kZM, NCOL, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(kZM, kD, 0, 1);
      constexpr uint32_t idesc3 = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t sg = smem_u32(s_g), se = smem_u32(s_e);
      const bool with_dq = p.part_dq != nullptr;
      for (int i = 0; i < ntiles; ++i) {
        const int stage = i & 1;
        const uint32_t sh = smem_u32(tiles + (size_t)stage * kZTile);
        mbar_wait_idle(&full[stage], (i >> 1) & 1);
        mbar_wait_idle(sempty, (i & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {       // S = h G^T, K = 256 features
          const uint64_t ad = umma_desc_sw128(sh + (k >> 2) * (kZM * 128) + (k & 3) * 32, 0, 1024);
          const uint64_t bd = umma_desc_sw128(sg + (k >> 2) * (NCOL * 128) + (k & 3) * 32, 0, 1024);
          umma_f16(tm_s, ad, bd, idesc1, k != 0);
        }
        umma_commit(sfull);
        mbar_wait_idle(efull, i & 1);
        mbar_wait_idle(dempty, (i & 1) ^ 1);
        tc_fence_after();
        if (with_dq) {                            // before MMA2: its commit (dfull) lets epilogue 2 overwrite the h tile
          const uint32_t sebox = se + (dq_col >> 6) * (kZM * 128);
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int k = 0; k < kZM / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(sh + hf * 2 * (kZM * 128) + k * 2048, kZM * 128, 1024);
              const uint64_t bd = umma_desc_sw128(sebox + k * 2048, kZM * 128, 1024);
              umma_f16(tm_q + hf * 64, ad, bd, idesc3, (i | k) != 0);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < NCOL / 16; ++k) {     // dh = E G, K = NCOL stacked prototype rows
          const uint64_t ad = umma_desc_sw128(se + (k >> 2) * (kZM * 128) + (k & 3) * 32, 0, 1024);
          // G as an MN-major B operand: 64-feature boxes are NCOL*128 bytes apart, 8 k-rows are 1024 bytes apart
          const uint64_t bd = umma_desc_sw128(sg + k * 2048, NCOL * 128, 1024);
          umma_f16(tm_d, ad, bd, idesc2, k != 0);
        }
        umma_commit(dfull);
      }
    }
  } else {
    // ------------------------------ epilogue warps ------------------------------
    const int q = warp & 3, hf = warp >> 2;
    const int n = q * 32 + lane;                   // row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const bool nm = p.relu_mask == 0;
    const float keep = p.keep_scale;
    float dbacc = 0.f;                             // column sum for feature f = threadIdx.x
    for (int i = 0; i < ntiles; ++i) {
      const int stage = i & 1;
      uint8_t* tile = tiles + (size_t)stage * kZTile;
      const int tile_row0 = row_begin + (t0 + i) * kZM;
      const bool row_ok = tile_row0 + n < row_end;
      // ---- epilogue 1: S -> E ----
      mbar_wait(sfull, i & 1);
      tc_fence_after();
#pragma unroll
      for (int c8 = 0; c8 < NS / 8; ++c8) {
        const int slot0 = hf * NS + c8 * 8;        // prototype slots slot0..slot0+7 (never straddle a block: PP % 8 == 0)
        const int blk = slot0 / PP, p0 = slot0 % PP;
        const int cS = blk * 2 * PP + p0, cA = cS + PP;
        uint32_t sv[8], av[8];
        tmem_ld8(tm_s + lane_addr + cS, sv);
        tmem_ld8(tm_s + lane_addr + cA, av);
        tmem_ld_wait();
        float a[8], ds[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float lse = s_lse[slot0 + e], dl = s_delta[slot0 + e];
          a[e] = row_ok ? exp2f((__uint_as_float(sv[e]) - lse) * 1.4426950408889634f) : 0.f;
          ds[e] = a[e] * (__uint_as_float(av[e]) - dl);
        }
        *reinterpret_cast<uint4*>(s_e + box128_off(n, cS)) =
            make_uint4(pack_bf16x2(ds[0], ds[1]), pack_bf16x2(ds[2], ds[3]), pack_bf16x2(ds[4], ds[5]), pack_bf16x2(ds[6], ds[7]));
        *reinterpret_cast<uint4*>(s_e + box128_off(n, cA)) =
            make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
      }
      tc_fence_before();
      fence_proxy_async_smem();                    // E is read by MMA2 through the async proxy
      __syncwarp();
      if (lane == 0) { mbar_arrive(sempty); mbar_arrive(efull); }
      // ---- epilogue 2: dh -> dz (in place over the h tile) ----
      mbar_wait(dfull, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int f0 = hf * 128 + cc * 32;
        uint32_t v[32];
        tmem_ld32(tm_d + lane_addr + f0, v);
        tmem_ld_wait();
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          uint4* ph = reinterpret_cast<uint4*>(tile + box128_off(n, f0 + g8 * 8));
          const uint4 hv = *ph;
          const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float z0 = (nm || bf16lo(hw[e]) > 0.f) ? __uint_as_float(v[g8 * 8 + 2 * e]) * keep : 0.f;
            const float z1 = (nm || bf16hi(hw[e]) > 0.f) ? __uint_as_float(v[g8 * 8 + 2 * e + 1]) * keep : 0.f;
            o[e] = pack_bf16x2(z0, z1);
          }
          *ph = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dempty);
      bar_sync(1, 256);                            // the whole dz tile is in shared memory
      const int nvalid = min(kZM, row_end - tile_row0);
      for (int it = threadIdx.x; it < nvalid * 32; it += 256) {
        const int r = it >> 5, c = (it & 31) << 3;
        *reinterpret_cast<uint4*>(p.dz + (size_t)(tile_row0 + r) * kD + c) = *reinterpret_cast<const uint4*>(tile + box128_off(r, c));
      }
      if (p.part_db) {
        const int f = threadIdx.x;
        float acc = 0.f;
        for (int r = 0; r < nvalid; ++r) acc += __bfloat162float(*reinterpret_cast<const bf16*>(tile + box128_off(r, f)));
        dbacc += acc;
      }
      fence_proxy_async_smem();                    // generic-proxy writes of this stage before the next TMA load into it
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
    if (p.part_db) p.part_db[((size_t)b * p.nsplit + split) * kD + threadIdx.x] = dbacc;
    if (p.part_dq) {                               // every MMA of the CTA has retired (dfull of the last tile was waited for)
      float* out_dq = p.part_dq + ((size_t)b * p.nsplit + split) * PP * kD;
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < PP; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tm_q + lane_addr + hf * 64 + (dq_col & 63) + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) out_dq[(size_t)(c0 + e) * kD + hf * 128 + n] = __uint_as_float(v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dz rows past cu[B] belong to no bag (a packed buffer sized for the worst case): the dz kernel never visits them,
// but dW1 = dz^T x sums over every row of the buffer, so they must be zero rather than whatever the allocator left
__global__ void zero_tail_rows_kernel(bf16* __restrict__ dz, const int* __restrict__ cu, int B, int total_rows) {
  const size_t begin = (size_t)__ldg(cu + B) * (kD / 8), end = (size_t)total_rows * (kD / 8);      // in uint4 units
  for (size_t i = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (size_t)gridDim.x * blockDim.x)
    reinterpret_cast<uint4*>(dz)[i] = make_uint4(0u, 0u, 0u, 0u);
}

int pad_protos(int P) { return P <= 16 ? 16 : (P <= 32 ? 32 : 64); }

template <int PP, int STAGES>
constexpr size_t fwd_smem() {
  return 1024 + (size_t)STAGES * kTileBytes + PP * 512 + kTM * (PP * 2 + 16) + 5 * PP * 4 + 2 * STAGES * 8 + 64;
}
template <int PP, int NB, int STAGES>
constexpr size_t bwd_smem() {
  size_t e = (size_t)kTM * (2 * NB * PP * 2 + 16);
  if (e < 4 * kD * 4) e = 4 * kD * 4;
  return 1024 + (size_t)STAGES * kTileBytes + 2 * NB * PP * 512 + e + 2 * NB * PP * 4 + 2 * STAGES * 8 + 64;
}

}  // namespace

// Split every bag into `nsplit` runs of `tiles_per_split` tiles so that the B * nsplit CTAs fill whole waves of
// the `slots` CTAs the GPU holds at once: cost = waves * (tiles per CTA + 1 tile-equivalent of prologue).  Rounding
// the split count up to "at least two waves" (the first version) gave 608 CTAs on 296 slots for 32 bags: a third
// wave of 16 CTAs, 1.5x the time of two full waves.
void best_split(int tiles, int B, int slots, int min_tiles, int max_split, int* nsplit, int* tiles_per_split) {
  const int cap = max(1, min(tiles / max(1, min_tiles), max_split));
  long best_cost = -1;
  int best_ns = 1, best_tps = tiles;
  for (int ns = 1; ns <= cap; ++ns) {
    const int tps = (tiles + ns - 1) / ns;
    const int real_ns = (tiles + tps - 1) / tps;
    const long waves = ((long)B * real_ns + slots - 1) / slots;
    const long cost = waves * (tps + 1);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_ns = real_ns; best_tps = tps; }
  }
  *nsplit = best_ns;
  *tiles_per_split = best_tps;
}

namespace {

// forward / dq: 64-row tiles, two CTAs resident per SM up to 32 prototypes (one above), at least 8 tiles (256 KB of h) per CTA
void split_plan(int max_len, int B, int P, int* nsplit, int* tiles_per_split) {
  const int tiles = max(1, (max_len + kTM - 1) / kTM);
  best_split(tiles, B, (P <= 32 ? 2 : 1) * imp_num_sms(), 8, 256, nsplit, tiles_per_split);
}

template <int PP, int STAGES>
int run_fwd(const CUtensorMap& tm, const PoolFwdParams& p, int B, cudaStream_t st) {
  constexpr size_t smem = fwd_smem<PP, STAGES>();
  { const int rc_ = imp_ensure_smem((const void*)pool_fwd_kernel<PP, STAGES>, smem); if (rc_) return rc_; }
  IMP_LAUNCH("pool_fwd", st, pool_fwd_kernel<PP, STAGES><<<dim3(p.nsplit, B), kThreads, smem, st>>>(tm, p));
  return IMP_OK;
}
template <int PP, int NB, int STAGES>
int run_bwd(const CUtensorMap& tm, const PoolBwdParams& p, int B, cudaStream_t st) {
  constexpr size_t smem = bwd_smem<PP, NB, STAGES>();
  static_assert(smem <= 227 * 1024, "pool_bwd shared memory");
  { const int rc_ = imp_ensure_smem((const void*)pool_bwd_kernel<PP, NB, STAGES>, smem); if (rc_) return rc_; }
  IMP_LAUNCH("pool_bwd_dq", st, pool_bwd_kernel<PP, NB, STAGES><<<dim3(p.nsplit, B), kThreads, smem, st>>>(tm, p));
  return IMP_OK;
}

template <int PP, int NB>
int run_dz(const CUtensorMap& tm, const DzParams& p, int B, cudaStream_t st) {
  constexpr size_t smem = dz_smem<PP, NB>();
  static_assert(smem <= 227 * 1024, "pool_bwd_dz shared memory");
  { const int rc_ = imp_ensure_smem((const void*)pool_bwd_dz_kernel<PP, NB>, smem); if (rc_) return rc_; }
  IMP_LAUNCH("pool_bwd_dz", st, pool_bwd_dz_kernel<PP, NB><<<dim3(p.nsplit, B), kZThreads, smem, st>>>(tm, p));
  return IMP_OK;
}

// dz: 128-row tiles, one CTA per SM, at least 4 tiles (256 KB of h) per CTA
void dz_split_plan(int max_len, int B, int* nsplit, int* tiles_per_split) {
  const int tiles = max(1, (max_len + kZM - 1) / kZM);
  best_split(tiles, B, imp_num_sms(), 4, 128, nsplit, tiles_per_split);
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
// pool_tc.cu: the tcgen05 forward / dq~ kernels (128-row tiles, prototypes padded to 32 or 64)
int pool_tc_pad(int P);
void pool_tc_split_plan(int max_len, int B, int* nsplit, int* tiles_per_split);
int launch_pool_tc_fwd(const bf16* h, int total_rows, const int* cu, int B, const float* qt, long long qt_stride, int P,
                       int nsplit, int tiles_per_split, float* part_acc, float* part_ml, cudaStream_t st);
int launch_pool_tc_dq(const bf16* h, int total_rows, const int* cu, int B, const float* qt, long long qt_stride,
                      const float* dpool, long long dpool_stride, const float* lse, const float* delta, int P, int nsplit,
                      int tiles_per_split, float* part_dq, cudaStream_t st);
static bool use_legacy_pool() {      // bring-up switch: IMP_POOL_MMASYNC=1 selects the round-1 mma.sync kernels
  static const bool v = []() { const char* e = getenv("IMP_POOL_MMASYNC"); return e && atoi(e) != 0; }();
  return v;
}

size_t pool_fwd_workspace_bytes(int B, int max_len, int P) {
  int ns, tps;
  split_plan(max_len, B, P, &ns, &tps);
  const int PP = pad_protos(P);
  size_t legacy = ((size_t)B * ns * PP * kD + (size_t)B * ns * 2 * PP) * sizeof(float);
  pool_tc_split_plan(max_len, B, &ns, &tps);
  const int PT = pool_tc_pad(P);
  size_t tc = ((size_t)B * ns * PT * kD + (size_t)B * ns * 2 * PT) * sizeof(float);
  return legacy > tc ? legacy : tc;
}

int launch_pool_fwd(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* qt,
                    long long qt_stride, int P, float* workspace, float* pooled, float* lse, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  if (P <= 0 || P > 64) IMP_FAIL(IMP_ERR_ARG, "pool_fwd: P=%d out of [1,64]", P);
  if (total_rows <= 0 || max_len <= 0) IMP_FAIL(IMP_ERR_ARG, "pool_fwd: empty input (rows=%d, max_len=%d)", total_rows, max_len);
  if (!use_legacy_pool()) {
    const int PT = pool_tc_pad(P);
    int ns, tps;
    pool_tc_split_plan(max_len, B, &ns, &tps);
    float* part_acc = workspace;
    float* part_ml = workspace + (size_t)B * ns * PT * kD;
    int rc = launch_pool_tc_fwd(h, total_rows, cu, B, qt, qt_stride, P, ns, tps, part_acc, part_ml, st);
    if (rc) return rc;
    IMP_LAUNCH("pool_merge", st, pool_merge_kernel<<<dim3(P, B), kD, 0, st>>>(part_acc, part_ml, pooled, lse, P, PT, ns));
    return IMP_OK;
  }
  const int PP = pad_protos(P);
  PoolFwdParams p;
  split_plan(max_len, B, P, &p.nsplit, &p.tiles_per_split);
  p.cu = cu; p.qt = qt; p.qt_stride = qt_stride; p.P = P;
  p.part_acc = workspace;
  p.part_ml = workspace + (size_t)B * p.nsplit * PP * kD;
  CUtensorMap tm;
  int rc = imp_make_tmap_2d(&tm, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kTM,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (PP == 16) rc = run_fwd<16, 2>(tm, p, B, st);
  else if (PP == 32) rc = run_fwd<32, 2>(tm, p, B, st);
  else rc = run_fwd<64, 4>(tm, p, B, st);
  if (rc) return rc;
  IMP_LAUNCH("pool_merge", st, pool_merge_kernel<<<dim3(P, B), kD, 0, st>>>(p.part_acc, p.part_ml, pooled, lse, P, PP, p.nsplit));
  return IMP_OK;
}

size_t pool_bwd_workspace_bytes(int B, int max_len, int P) {
  int ns, tps;
  split_plan(max_len, B, P, &ns, &tps);
  const int PP = pad_protos(P);
  int nz, tz;
  dz_split_plan(max_len, B, &nz, &tz);
  int nt, tt;
  pool_tc_split_plan(max_len, B, &nt, &tt);
  const size_t dq_part = std::max(std::max((size_t)B * ns * PP * kD, (size_t)B * nt * pool_tc_pad(P) * kD), (size_t)B * nz * PP * kD);
  return (dq_part + (size_t)B * nz * kD) * sizeof(float);
}

int launch_pool_bwd(const bf16* h, int total_rows, const int* cu, int B, int max_len, int nblocks,
                    const float* const* qt, const long long* qt_stride, const float* const* dpool,
                    const float* const* lse, const float* const* delta, int P, int dq_block, int relu_mask,
                    float keep_scale, float* workspace, float* dq, bf16* dz, float* db1, int db_accumulate, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  if (P <= 0 || P > 64) IMP_FAIL(IMP_ERR_ARG, "pool_bwd: P=%d out of [1,64]", P);
  if (nblocks < 1 || nblocks > 2) IMP_FAIL(IMP_ERR_ARG, "pool_bwd: nblocks=%d (1 or 2)", nblocks);
  if (dq_block < 0 || dq_block >= nblocks) IMP_FAIL(IMP_ERR_ARG, "pool_bwd: dq_block=%d", dq_block);
  if (total_rows <= 0 || max_len <= 0) IMP_FAIL(IMP_ERR_ARG, "pool_bwd: empty input");
  if (nblocks == 2 && P > 32) IMP_FAIL(IMP_ERR_ARG, "pool_bwd: two stacked blocks need P <= 32 (got %d)", P);
  const int PP = pad_protos(P);
  PoolBwdParams p;
  split_plan(max_len, B, P, &p.nsplit, &p.tiles_per_split);
  p.cu = cu; p.P = P; p.dq_block = dq_block;
  for (int k = 0; k < 2; ++k) {
    const int s = k < nblocks ? k : 0;
    p.qt[k] = qt[s]; p.qt_stride[k] = qt_stride[s];
    p.dpool[k] = dpool[s]; p.dpool_stride[k] = (long long)P * kD;
    p.lse[k] = lse[s]; p.delta[k] = delta[s];
  }
  p.part_dq = workspace;
  float* part_db = nullptr;
  CUtensorMap tm;
  int rc = imp_make_tmap_2d(&tm, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kTM,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  // (1) dq~ of the requested block: streaming kernel over that block alone (tcgen05; mma.sync behind the switch)
  int dq_ns = p.nsplit, dq_pp = PP;
  size_t dq_part_elems = (size_t)B * p.nsplit * PP * kD;
  {
    int nt, tt;
    pool_tc_split_plan(max_len, B, &nt, &tt);
    dq_part_elems = std::max(dq_part_elems, (size_t)B * nt * pool_tc_pad(P) * kD);
    int nz0, tz0;
    dz_split_plan(max_len, B, &nz0, &tz0);
    dq_part_elems = std::max(dq_part_elems, (size_t)B * nz0 * PP * kD);
    if (!use_legacy_pool()) { dq_ns = nt; dq_pp = pool_tc_pad(P); }
  }
  const bool fused_dq = dz != nullptr && !use_legacy_pool();     // the dz kernel forms dq~ of `dq_block` on the way
  if (fused_dq) {
    int nz0, tz0;
    dz_split_plan(max_len, B, &nz0, &tz0);
    dq_ns = nz0; dq_pp = PP;
  } else if (!use_legacy_pool()) {
    int nt, tt;
    pool_tc_split_plan(max_len, B, &nt, &tt);
    rc = launch_pool_tc_dq(h, total_rows, cu, B, qt[dq_block], qt_stride[dq_block], dpool[dq_block], (long long)P * kD,
                           lse[dq_block], delta[dq_block], P, nt, tt, p.part_dq, st);
    if (rc) return rc;
  } else {
    PoolBwdParams q1 = p;
    q1.qt[0] = qt[dq_block]; q1.qt_stride[0] = qt_stride[dq_block];
    q1.dpool[0] = dpool[dq_block]; q1.lse[0] = lse[dq_block]; q1.delta[0] = delta[dq_block];
    q1.dq_block = 0;
#define IMP_BWD(PPv, STv) rc = run_bwd<PPv, 1, STv>(tm, q1, B, st)
    if (PP == 16) IMP_BWD(16, 2); else if (PP == 32) IMP_BWD(32, 2); else IMP_BWD(64, 3);
#undef IMP_BWD
    if (rc) return rc;
  }
  // (2) dz (and db1): both blocks at once on tcgen05
  int nz = 0;
  if (dz) {
    DzParams z;
    dz_split_plan(max_len, B, &z.nsplit, &z.tiles_per_split);
    nz = z.nsplit;
    z.cu = cu; z.P = P; z.relu_mask = relu_mask; z.keep_scale = relu_mask ? keep_scale : 1.f; z.dz = dz;
    for (int k = 0; k < 2; ++k) {
      z.qt[k] = p.qt[k]; z.qt_stride[k] = p.qt_stride[k]; z.dpool[k] = p.dpool[k]; z.dpool_stride[k] = p.dpool_stride[k];
      z.lse[k] = p.lse[k]; z.delta[k] = p.delta[k];
    }
    z.part_db = db1 ? workspace + dq_part_elems : nullptr;
    z.part_dq = fused_dq ? p.part_dq : nullptr;
    z.dq_block = dq_block;
    CUtensorMap tmz;
    if ((rc = imp_make_tmap_2d(&tmz, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kZM,
                               CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if (PP == 16) rc = nblocks == 1 ? run_dz<16, 1>(tmz, z, B, st) : run_dz<16, 2>(tmz, z, B, st);
    else if (PP == 32) rc = nblocks == 1 ? run_dz<32, 1>(tmz, z, B, st) : run_dz<32, 2>(tmz, z, B, st);
    else rc = run_dz<64, 1>(tmz, z, B, st);
    if (rc) return rc;
    part_db = z.part_db;
    IMP_LAUNCH("zero_tail_rows", st, zero_tail_rows_kernel<<<imp_num_sms(), 256, 0, st>>>(dz, cu, B, total_rows));
  }
  IMP_LAUNCH("reduce_dq", st, reduce_dq_kernel<<<dim3(P, B), kD, 0, st>>>(p.part_dq, dq, P, dq_pp, dq_ns));
  if (part_db) {
    IMP_LAUNCH("reduce_db", st, reduce_db_kernel<<<1, kD, 0, st>>>(part_db, db1, B * nz, db_accumulate));
  }
  return IMP_OK;
}


// Cross-GPU merge of per-rank pooling results: rank r holds (pooled_r, lse_r) of its
// patch shard; as a partial state that is acc = pooled_r, m = lse_r, l = 1, so the same
// log-sum-exp merge kernel applies.  part_pooled (B, nsplit, P, 256), part_lse (B, nsplit, P).
namespace {
__global__ void lse_to_ml_kernel(const float* __restrict__ part_lse, float* __restrict__ ml, int P, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over (B*nsplit, P)
  if (i >= total) return;
  const int bs = i / P, pi = i % P;
  const float v = part_lse[i];
  ml[(size_t)bs * 2 * P + pi] = v;
  ml[(size_t)bs * 2 * P + P + pi] = (v == -INFINITY) ? 0.f : 1.f;    // an empty shard carries no mass
}
}  // namespace

int launch_lse_merge(const float* part_pooled, const float* part_lse, int B, int nsplit, int P, float* pooled,
                     float* lse, float* scratch, cudaStream_t st) {
  if (B <= 0 || nsplit <= 0 || P <= 0) IMP_FAIL(IMP_ERR_ARG, "lse_merge: bad shape (%d,%d,%d)", B, nsplit, P);
  const int total = B * nsplit * P;
  IMP_LAUNCH("lse_to_ml", st, lse_to_ml_kernel<<<(total + 255) / 256, 256, 0, st>>>(part_lse, scratch, P, total));
  IMP_LAUNCH("pool_merge", st, pool_merge_kernel<<<dim3(P, B), kD, 0, st>>>(part_pooled, scratch, pooled, lse, P, P, nsplit));
  return IMP_OK;
}
