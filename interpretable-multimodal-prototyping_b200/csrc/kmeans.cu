// A9  k-means prototype assignment (new functionality: the reference ships no k-means, SURVEY.md D1;
// the distance follows medmm/metrics/distance.py:46-61, ||a||^2 + ||b||^2 - 2 a.b, then argmin with the
// first index winning ties).
//
//   assign[n] = argmin_k ( ||x_n||^2 + ||mu_k||^2 - 2 x_n . mu_k )      x (N,D) fp32, mu (K,D) fp32, K <= 64
//
// Two paths, both fp32-accurate (exact agreement with the fp32 oracle is required, so nothing is computed in
// reduced precision):  (1) K <= 32: tcgen05 tensor cores on three-way bf16 splits of x and mu (every fp32 number is
// exactly hi + mid + lo in bf16; the nine partial products are accumulated in fp32 in TMEM), kmeans_assign_tc_kernel
// below -- the default for configs[4];  (2) K > 32 or large D: fp32 FMA on CUDA cores.  Algorithmically HBM-bound
// (N*D*4 bytes read once), but 2*K FLOP per byte put the fp32 pipe right at the roofline: the scalar FFMA rate
// (64 lanes/clk/SM) is too slow, so the inner product runs on packed FFMA2 (two centroids per instruction, x
// broadcast) with RPT rows per thread sharing every broadcast load of mu^T from shared memory.  One TMA producer warp streams [256*RPT rows][16 floats] tiles (64-byte swizzle);
// 8 consumer warps, thread = RPT rows.  The accumulation order over the features is the oracle's
// (ascending d, one fma per (row, centroid, d)), so distances are bit-identical to the scalar kernel.
// Also: the Lloyd update (sums and counts per centroid).
#include "common.cuh"
#include "launchers.h"
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int kKC = 16;                  // floats per k-chunk (one 64 B swizzle row)
constexpr int kStages = 2;
constexpr int kThreads = 256 + 32;

// acc (two packed fp32 accumulators, kept as one opaque 64-bit register pair) += x * (m0, m1)
__device__ __forceinline__ void km_fma2(uint64_t& acc, float x, float m0, float m1) {
  asm("{\n\t"
      ".reg .b64 xb, mb;\n\t"
      "mov.b64 xb, {%1, %1};\n\t"
      "mov.b64 mb, {%2, %3};\n\t"
      "fma.rn.f32x2 %0, xb, mb, %0;\n\t"
      "}"
      : "+l"(acc)
      : "f"(x), "f"(m0), "f"(m1));
}

struct KmParams {
  const float* mu;      // (K,D)
  int* assign;          // (N)
  float* best_dist;     // (N) or null
  int N, D, K, num_tiles;
};

template <int KP, int RPT>      // centroids padded to 8/16/32/64; rows per thread
__global__ void __launch_bounds__(kThreads, 1)
kmeans_assign_kernel(const __grid_constant__ CUtensorMap tm_x, const KmParams p) {
  constexpr int kTileRows = 256 * RPT;
  constexpr int kStageBytes = kTileRows * kKC * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* tiles = smem;
  float* s_mu = reinterpret_cast<float*>(tiles + kStages * kStageBytes);      // [D][KP] transposed
  float* s_m2 = s_mu + (size_t)p.D * KP;                                      // [KP]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_m2 + KP);
  uint64_t* empty = full + kStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = p.D / kKC;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < p.D * KP; i += kThreads) {
    const int k = i % KP, d = i / KP;
    s_mu[i] = k < p.K ? __ldg(p.mu + (size_t)k * p.D + d) : 0.f;
  }
  __syncthreads();
  if (threadIdx.x < KP) {
    float a = 0.f;
    if (threadIdx.x < p.K) for (int d = 0; d < p.D; ++d) { const float v = s_mu[d * KP + threadIdx.x]; a = fmaf(v, v, a); }
    s_m2[threadIdx.x] = a;
  }
  __syncthreads();

  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
        for (int c = 0; c < nchunks; ++c, ++it) {
          const int stage = it % kStages;
          mbar_wait_idle(&empty[stage], ((it / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          uint8_t* dst = tiles + (size_t)stage * kStageBytes;
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr)        // TMA boxes hold at most 256 rows
            tma_load_2d(dst + rr * (256 * kKC * 4), &tm_x, &full[stage], c * kKC, tile * kTileRows + rr * 256);
        }
    }
    return;
  }

  const int r = threadIdx.x;               // rows r, r+256, ... of the tile
  const uint32_t swz = (uint32_t)((r >> 1) & 3);       // 64-byte swizzle: 16-byte chunk ^= (row / 2) % 4
  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    uint64_t acc[RPT][KP / 2];             // packed pairs (centroid 2k, 2k+1)
    float xx[RPT];
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
      xx[rr] = 0.f;
#pragma unroll
      for (int k = 0; k < KP / 2; ++k) acc[rr][k] = 0ull;
    }
    for (int c = 0; c < nchunks; ++c, ++it) {
      const int stage = it % kStages;
      mbar_wait(&full[stage], (it / kStages) & 1);
      const uint8_t* trow = tiles + (size_t)stage * kStageBytes + r * (kKC * 4);
      const float* mu_c = s_mu + (size_t)c * kKC * KP;
#pragma unroll
      for (int q = 0; q < kKC / 4; ++q) {
        float xs[RPT][4];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) {
          const float4 xv = *reinterpret_cast<const float4*>(trow + rr * (256 * kKC * 4) + (((uint32_t)q ^ swz) << 4));
          xs[rr][0] = xv.x; xs[rr][1] = xv.y; xs[rr][2] = xv.z; xs[rr][3] = xv.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr) xx[rr] = fmaf(xs[rr][e], xs[rr][e], xx[rr]);
          const float4* m4 = reinterpret_cast<const float4*>(mu_c + (q * 4 + e) * KP);
#pragma unroll
          for (int k4 = 0; k4 < KP / 4; ++k4) {
            const float4 m = m4[k4];                       // warp-broadcast: 4 centroids of feature d
#pragma unroll
            for (int rr = 0; rr < RPT; ++rr) {
              km_fma2(acc[rr][2 * k4], xs[rr][e], m.x, m.y);
              km_fma2(acc[rr][2 * k4 + 1], xs[rr][e], m.z, m.w);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
#pragma unroll
    for (int rr = 0; rr < RPT; ++rr) {
      const int row = tile * kTileRows + rr * 256 + r;
      if (row < p.N) {
        float best = INFINITY;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const float a = __uint_as_float((k & 1) ? (uint32_t)(acc[rr][k >> 1] >> 32) : (uint32_t)acc[rr][k >> 1]);
          const float dist = fmaf(-2.f, a, xx[rr] + s_m2[k]);      // (xx + mm) - 2 x.mu, distance.py:55-60
          if (k < p.K && dist < best) { best = dist; arg = k; }   // strict <: first index wins ties
        }
        p.assign[row] = arg;
        if (p.best_dist) p.best_dist[row] = best;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Tensor-core path (K <= 32, D % 64 == 0, 3 K D 2 bytes of centroid parts <= 96 KB): fp32-accurate inner products
// from bf16 MMAs.  Every fp32 number is exactly the sum of three bf16 numbers (8 + 8 + 8 significant bits):
// x = xh + xm + xl, mu = mh + mm + ml (split by truncation, exact).  The centroid parts are stacked along N
// ([mh; mm; ml], N = 3 KP), so one tcgen05 MMA per x part yields x.mh, x.mm and x.ml in three column groups of the
// fp32 TMEM accumulator: all nine partial products, exact in bf16 x bf16 -> fp32, summed by the tensor core; the
// epilogue adds the three groups, small ones first.  The CUDA cores only split the stream
// (4.5 instructions per element: two masks, two packed subtractions, three byte permutes per pair), so the kernel
// is bound by HBM instead of by the fp32 pipe.
//   warps 0-7  : split the fp32 TMA boxes [128 rows][32 features] into three bf16 K-major operand tiles (the two
//                feature halves of a 64-wide chunk are pipelined against the MMAs of the other half), |x|^2,
//                and (warps 0-3, thread = row) the arg-min epilogue from TMEM
//   warp 8     : TMA producer (4 box slots);   warp 9 : MMA issuer (3 x parts x 2 k-steps per half chunk, N = 3 KP)
// ------------------------------------------------------------------------------------------
constexpr int kTcRows = 128;
constexpr int kTcKC = 64;                                  // features per chunk
constexpr int kTcXStage = kTcRows * kTcKC * 4;             // 32 KB fp32: two [128][32] boxes
constexpr int kTcPart = kTcRows * kTcKC * 2;               // 16 KB bf16 operand tile
constexpr int kTcThreads = 10 * 32;

struct KmTcParams {
  const float* mu;
  int* assign;
  float* best_dist;
  int N, D, K, num_tiles;
};

// Exact three-way split by truncation: h = the top 16 bits of v (a bf16 number), r = v - h (exact), m = the top 16
// bits of r, l = r - m (exact, at most 8 significant bits, i.e. again a bf16 number):  v == h + m + l.
__device__ __forceinline__ void split3(float v, float& h, float& m, float& l) {
  h = __uint_as_float(__float_as_uint(v) & 0xffff0000u);
  const float r1 = v - h;
  m = __uint_as_float(__float_as_uint(r1) & 0xffff0000u);
  l = r1 - m;
}
// the same for two numbers at once: the subtractions are one packed FFMA2 each, the three bf16x2 words come out
// of byte permutes of the upper halves (no rounding is involved anywhere)
__device__ __forceinline__ void split3x2(float v0, float v1, uint32_t& wh, uint32_t& wm, uint32_t& wl) {
  const uint32_t h0 = __float_as_uint(v0) & 0xffff0000u, h1 = __float_as_uint(v1) & 0xffff0000u;
  uint64_t r;
  {
    const float2 hv = make_float2(__uint_as_float(h0), __uint_as_float(h1)), vv = make_float2(v0, v1), neg = make_float2(-1.f, -1.f);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<const uint64_t*>(&hv)), "l"(*reinterpret_cast<const uint64_t*>(&neg)),
        "l"(*reinterpret_cast<const uint64_t*>(&vv)));
  }
  const uint32_t r0 = (uint32_t)r, r1 = (uint32_t)(r >> 32);
  const uint32_t m0 = r0 & 0xffff0000u, m1 = r1 & 0xffff0000u;
  uint64_t l;
  {
    const float2 mv = make_float2(__uint_as_float(m0), __uint_as_float(m1)), neg = make_float2(-1.f, -1.f);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(l) : "l"(*reinterpret_cast<const uint64_t*>(&mv)), "l"(*reinterpret_cast<const uint64_t*>(&neg)), "l"(r));
  }
  wh = __byte_perm(h0, h1, 0x7632);                        // (bf16(v0) | bf16(v1) << 16): upper halves of both words
  wm = __byte_perm(m0, m1, 0x7632);
  wl = __byte_perm((uint32_t)l, (uint32_t)(l >> 32), 0x7632);
}

template <int KP>
__global__ void __launch_bounds__(kTcThreads, 1)
kmeans_assign_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const KmTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nchunks = p.D / kTcKC;
  constexpr int kMuBox = 3 * KP * 128;                     // per chunk: [mh; mm; ml] = 3 KP rows of 64 features
  uint8_t* s_x = smem;                                     // 2 stages of the fp32 tile
  uint8_t* s_parts = s_x + 2 * kTcXStage;                  // xh | xm | xl operand tiles
  uint8_t* s_mu = s_parts + 3 * kTcPart;                   // mh | mm | ml
  float* s_m2 = reinterpret_cast<float*>(s_mu + (size_t)nchunks * kMuBox);
  float* s_xx = s_m2 + KP;                                 // [128] upper-half partial of |x|^2
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_xx + kTcRows);
  uint64_t* xfull = bars;          // [4] TMA -> splitters, one slot = one [128][32 floats] box = half a chunk
  uint64_t* xempty = bars + 4;     // [4] 8 warps -> TMA
  uint64_t* pfull = bars + 8;      // [2] 8 warps -> MMA: operand tiles of feature half 0 / 1 written
  uint64_t* pempty = bars + 10;    // [2] MMA commit -> splitters
  uint64_t* tfull = bars + 12;     // [2] MMA -> epilogue
  uint64_t* tempty = bars + 14;    // [2] 4 warps -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < 4; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(&pfull[i], 8); mbar_init(&pempty[i], 1); mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 9) tmem_alloc(tmem_slot, 256);
  // centroid parts, K-major 128-byte-swizzled boxes [KP rows][64 features] per chunk; padded centroids are zero
  for (int i = threadIdx.x; i < KP * (p.D / 8); i += kTcThreads) {
    const int k = i / (p.D / 8), d0 = (i % (p.D / 8)) * 8;
    uint32_t w[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float h[2], m[2], l[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float v = k < p.K ? __ldg(p.mu + (size_t)k * p.D + d0 + 2 * e + j) : 0.f;
        split3(v, h[j], m[j], l[j]);
      }
      w[0][e] = pack_bf16x2(h[0], h[1]); w[1][e] = pack_bf16x2(m[0], m[1]); w[2][e] = pack_bf16x2(l[0], l[1]);
    }
#pragma unroll
    for (int part = 0; part < 3; ++part) {
      const int row = part * KP + k;                       // KP is a multiple of 8: the swizzle phase of a row is k & 7
      const uint32_t off = (uint32_t)((d0 >> 6) * kMuBox + row * 128 + ((((d0 & 63) >> 3) ^ (row & 7)) << 4));
      *reinterpret_cast<uint4*>(s_mu + off) = make_uint4(w[part][0], w[part][1], w[part][2], w[part][3]);
    }
  }
  if (threadIdx.x < KP) {
    float a = 0.f;
    if (threadIdx.x < p.K) for (int d = 0; d < p.D; ++d) { const float v = __ldg(p.mu + (size_t)threadIdx.x * p.D + d); a = fmaf(v, v, a); }
    s_m2[threadIdx.x] = a;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      int hh = 0;                                          // half-chunk counter: ring slot = hh & 3
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x)
        for (int c2 = 0; c2 < 2 * nchunks; ++c2, ++hh) {
          const int s = hh & 3;
          mbar_wait_idle(&xempty[s], ((hh >> 2) & 1) ^ 1);
          mbar_arrive_expect_tx(&xfull[s], kTcRows * 128);
          tma_load_2d(s_x + (size_t)s * (kTcRows * 128), &tm_x, &xfull[s], c2 * 32, tile * kTcRows);
        }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcRows, 3 * KP, 0, 0);
      const uint32_t sp = smem_u32(s_parts), sm = smem_u32(s_mu);
      int g = 0, ti = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
        const int acc = ti & 1;
        mbar_wait_idle(&tempty[acc], ((ti >> 1) & 1) ^ 1);
        for (int c = 0; c < nchunks; ++c, ++g) {
          const uint32_t sb = sm + c * kMuBox;
#pragma unroll
          for (int half = 0; half < 2; ++half) {           // the two 32-feature halves of the chunk are pipelined
            mbar_wait_idle(&pfull[half], g & 1);
            tc_fence_after();
#pragma unroll
            for (int t = 2; t >= 0; --t) {                 // x parts, smallest first: xl, xm, xh
              const uint32_t sa = sp + t * kTcPart;
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int ks = half * 2 + k2;
                umma_f16(tmem_base + acc * 128, umma_desc_sw128(sa + ks * 32, 0, 1024), umma_desc_sw128(sb + ks * 32, 0, 1024),
                         idesc, (c != 0) || (half != 0) || (t != 2) || (k2 != 0));
              }
            }
            umma_commit(&pempty[half]);
          }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ---------------- splitters: thread = (row, 32-feature half of the chunk) ----------------
    const int r = threadIdx.x & 127, hf = threadIdx.x >> 7;   // hf: which 16 of the 32 features of a half-chunk
    const bool want_dist = p.best_dist != nullptr;         // |x|^2 is the same for every centroid: only the distance output needs it
    int g = 0, ti = 0, hh = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ti) {
      float xx = 0.f;
      for (int c = 0; c < nchunks; ++c, ++g) {
#pragma unroll
        for (int half = 0; half < 2; ++half, ++hh) {
          const int s = hh & 3;
          mbar_wait(&xfull[s], (hh >> 2) & 1);
          const uint8_t* xrow = s_x + (size_t)s * (kTcRows * 128) + r * 128;
          float v[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 q4 = *reinterpret_cast<const float4*>(xrow + (((hf * 4 + j) ^ (r & 7)) << 4));
            v[4 * j] = q4.x; v[4 * j + 1] = q4.y; v[4 * j + 2] = q4.z; v[4 * j + 3] = q4.w;
          }
          mbar_wait(&pempty[half], (g & 1) ^ 1);           // the MMAs of the previous chunk have read this half of the tiles
#pragma unroll
          for (int j = 0; j < 2; ++j) {                    // 8 features -> one 16-byte chunk per part
            uint32_t w[3][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split3x2(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], w[0][e], w[1][e], w[2][e]);
              if (want_dist) {
                xx = fmaf(v[8 * j + 2 * e], v[8 * j + 2 * e], xx);
                xx = fmaf(v[8 * j + 2 * e + 1], v[8 * j + 2 * e + 1], xx);
              }
            }
            const uint32_t off = (uint32_t)(r * 128 + (((half * 4 + hf * 2 + j) ^ (r & 7)) << 4));
#pragma unroll
            for (int part = 0; part < 3; ++part)
              *reinterpret_cast<uint4*>(s_parts + part * kTcPart + off) = make_uint4(w[part][0], w[part][1], w[part][2], w[part][3]);
          }
          fence_proxy_async_smem();
          __syncwarp();
          // the fp32 slot is released only now: releasing it right after the loads let the next TMA write race with them
          if (lane == 0) { mbar_arrive(&pfull[half]); mbar_arrive(&xempty[s]); }
        }
      }
      // ---------------- epilogue: thread = row (warps 0-3) ----------------
      if (p.best_dist) {
        if (hf == 1) s_xx[r] = xx;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (hf == 0) xx += s_xx[r];
      }
      if (hf == 0) {
        const int acc = ti & 1;
        mbar_wait(&tfull[acc], (ti >> 1) & 1);
        tc_fence_after();
        float best = INFINITY;
        int arg = 0;
#pragma unroll
        for (int c16 = 0; c16 < KP / 16; ++c16) {
          uint32_t dh[16], dm[16], dl[16];                  // x.mh, x.mm, x.ml of 16 centroids
          const uint32_t ta = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 128 + c16 * 16;
          tmem_ld16(ta, dh);
          tmem_ld16(ta + KP, dm);
          tmem_ld16(ta + 2 * KP, dl);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int kk = c16 * 16 + k;
            const float dot = (__uint_as_float(dl[k]) + __uint_as_float(dm[k])) + __uint_as_float(dh[k]);
            const float dist = fmaf(-2.f, dot, s_m2[kk]);                        // |mu|^2 - 2 x.mu  (+ |x|^2: same for all k)
            if (kk < p.K && dist < best) { best = dist; arg = kk; }              // strict <: first index wins ties
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const int row = tile * kTcRows + r;
        if (row < p.N) {
          p.assign[row] = arg;
          if (p.best_dist) p.best_dist[row] = best + xx;                         // distance.py:55-60
        }
      }
      if (p.best_dist) asm volatile("bar.sync 1, 256;" ::: "memory");          // s_xx is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Lloyd update: sums[k][d] += x[n][d], counts[k] += 1 for assign[n] == k.  One warp per row,
// CTA-private accumulators in shared memory, one global atomic per (k,d) per CTA.
__global__ void __launch_bounds__(256)
kmeans_update_kernel(const float* __restrict__ x, const int* __restrict__ assign, float* __restrict__ sums,
                     int* __restrict__ counts, int N, int D, int K) {
  extern __shared__ float s_sum[];            // [K][D] + counts
  int* s_cnt = reinterpret_cast<int*>(s_sum + (size_t)K * D);
  for (int i = threadIdx.x; i < K * D; i += blockDim.x) s_sum[i] = 0.f;
  for (int i = threadIdx.x; i < K; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = blockIdx.x * 8 + warp; n < N; n += gridDim.x * 8) {
    const int k = __ldg(assign + n);
    const float* row = x + (size_t)n * D;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(row + d));
      float* dst = s_sum + (size_t)k * D + d;
      atomicAdd(dst, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
    }
    if (lane == 0) atomicAdd(s_cnt + k, 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * D; i += blockDim.x)
    if (s_sum[i] != 0.f) atomicAdd(sums + i, s_sum[i]);
  for (int i = threadIdx.x; i < K; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(counts + i, s_cnt[i]);
}

template <int KP, int RPT>
int run_assign(const CUtensorMap& tm, KmParams p, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)kStages * 256 * RPT * kKC * 4 + (size_t)p.D * KP * 4 + KP * 4 + 2 * kStages * 8 + 64;
  if (smem > 227 * 1024) IMP_FAIL(IMP_ERR_ARG, "kmeans_assign: D=%d with %d centroids exceeds shared memory", p.D, p.K);
  { const int rc_ = imp_ensure_smem((const void*)kmeans_assign_kernel<KP, RPT>, smem); if (rc_) return rc_; }
  p.num_tiles = (p.N + 256 * RPT - 1) / (256 * RPT);
  const int grid = std::min(p.num_tiles, imp_num_sms());
  IMP_LAUNCH("kmeans_assign", st, kmeans_assign_kernel<KP, RPT><<<grid, kThreads, smem, st>>>(tm, p));
  return IMP_OK;
}

}  // namespace

int launch_kmeans_assign(const float* x, const float* mu, int N, int D, int K, int* assign, float* best_dist,
                         cudaStream_t st) {
  if (N <= 0) return IMP_OK;
  if (K < 1 || K > 64) IMP_FAIL(IMP_ERR_ARG, "kmeans_assign: K=%d out of [1,64]", K);
  if (D <= 0 || D % 32 != 0) IMP_FAIL(IMP_ERR_ARG, "kmeans_assign: D=%d must be a positive multiple of 32", D);
  CUtensorMap tm;
  int rc;
  // tensor-core path: K <= 32 centroids whose three bf16 parts fit next to the stream buffers
  const int KPt = K <= 16 ? 16 : 32;
  static const bool force_fma = []() { const char* e = getenv("IMP_KMEANS_FFMA"); return e && atoi(e) != 0; }();   // tests: compare both paths
  if (!force_fma && K <= 32 && D % kTcKC == 0 && (size_t)3 * KPt * D * 2 <= 96 * 1024) {
    if ((rc = imp_make_tmap_2d(&tm, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, D, N, (uint64_t)D * 4, 32, kTcRows,
                               CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    KmTcParams q;
    q.mu = mu; q.assign = assign; q.best_dist = best_dist; q.N = N; q.D = D; q.K = K;
    q.num_tiles = (N + kTcRows - 1) / kTcRows;
    const size_t smem = 1024 + 2 * kTcXStage + 3 * kTcPart + (size_t)3 * KPt * D * 2 + (KPt + kTcRows) * 4 + 256;
    const int grid = std::min(q.num_tiles, imp_num_sms());
    if (KPt == 16) {
      { const int rc_ = imp_ensure_smem((const void*)kmeans_assign_tc_kernel<16>, smem); if (rc_) return rc_; }
      IMP_LAUNCH("kmeans_assign", st, kmeans_assign_tc_kernel<16><<<grid, kTcThreads, smem, st>>>(tm, q));
    } else {
      { const int rc_ = imp_ensure_smem((const void*)kmeans_assign_tc_kernel<32>, smem); if (rc_) return rc_; }
      IMP_LAUNCH("kmeans_assign", st, kmeans_assign_tc_kernel<32><<<grid, kTcThreads, smem, st>>>(tm, q));
    }
    return IMP_OK;
  }
  rc = imp_make_tmap_2d(&tm, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, D, N, (uint64_t)D * 4, kKC, 256,
                        CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  KmParams p;
  p.mu = mu; p.assign = assign; p.best_dist = best_dist; p.N = N; p.D = D; p.K = K;
  p.num_tiles = 0;
  // rows per thread: as many as registers (RPT * K accumulators) and shared memory (mu^T + 2 stages) allow
  const bool big = (size_t)D * 4 * 32 > 96 * 1024;      // mu^T of 32 centroids above 96 KB: halve the x stages
  if (K <= 8) return run_assign<8, 4>(tm, p, st);
  if (K <= 16) return run_assign<16, 4>(tm, p, st);
  if (K <= 32) return big ? run_assign<32, 2>(tm, p, st) : run_assign<32, 3>(tm, p, st);   // 96 accumulator registers
  return run_assign<64, 1>(tm, p, st);
}

int launch_kmeans_update(const float* x, const int* assign, int N, int D, int K, float* sums, int* counts,
                         cudaStream_t st) {
  if (N <= 0) return IMP_OK;
  if (D % 4 != 0) IMP_FAIL(IMP_ERR_ARG, "kmeans_update: D=%d must be a multiple of 4", D);
  const size_t smem = (size_t)K * D * 4 + (size_t)K * 4;
  if (smem > 200 * 1024) IMP_FAIL(IMP_ERR_ARG, "kmeans_update: K*D too large for shared accumulators");
  { const int rc_ = imp_ensure_smem((const void*)kmeans_update_kernel, smem); if (rc_) return rc_; }
  const int grid = std::min((N + 63) / 64, 2 * imp_num_sms());
  IMP_LAUNCH("kmeans_update", st, kmeans_update_kernel<<<grid, 256, smem, st>>>(x, assign, sums, counts, N, D, K));
  return IMP_OK;
}
