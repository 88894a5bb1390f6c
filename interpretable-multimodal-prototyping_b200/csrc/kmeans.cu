// A9  k-means prototype assignment (new functionality: the reference ships no k-means, SURVEY.md D1;
// the distance follows medmm/metrics/distance.py:46-61, ||a||^2 + ||b||^2 - 2 a.b, then argmin with the
// first index winning ties).
//
//   assign[n] = argmin_k ( ||x_n||^2 + ||mu_k||^2 - 2 x_n . mu_k )      x (N,D) fp32, mu (K,D) fp32, K <= 64
//
// fp32 FMA on CUDA cores (exact agreement with the fp32 oracle is required, so no reduced-precision
// tensor-core path).  Algorithmically HBM-bound (N*D*4 bytes read once), but 2*K FLOP per byte put the fp32 pipe
// right at the roofline: the scalar FFMA rate (64 lanes/clk/SM) is too slow, so the inner product runs on packed
// FFMA2 (two centroids per instruction, x broadcast) with RPT rows per thread sharing every broadcast load of
// mu^T from shared memory.  One TMA producer warp streams [256*RPT rows][16 floats] tiles (64-byte swizzle);
// 8 consumer warps, thread = RPT rows.  The accumulation order over the features is the oracle's
// (ascending d, one fma per (row, centroid, d)), so distances are bit-identical to the scalar kernel.
// Also: the Lloyd update (sums and counts per centroid).
#include "common.cuh"
#include "launchers.h"
#include <algorithm>

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
  static size_t attr = 0;
  if (smem > attr) {
    IMP_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<KP, RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
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
  int rc = imp_make_tmap_2d(&tm, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, D, N, (uint64_t)D * 4, kKC, 256,
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
  static size_t attr = 0;
  if (smem > attr) {
    IMP_CUDA(cudaFuncSetAttribute(kmeans_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int grid = std::min((N + 63) / 64, 2 * imp_num_sms());
  IMP_LAUNCH("kmeans_update", st, kmeans_update_kernel<<<grid, 256, smem, st>>>(x, assign, sums, counts, N, D, K));
  return IMP_OK;
}
