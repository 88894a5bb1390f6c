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
// The forward and the dq~-only pass are in pool_tc.cu (tcgen05: scores, weights and the P x 256 accumulation all on the
// tensor cores, h^T read as an MN-major operand).  This file holds the dz pass -- the gradient that flows back into
// path_net, a pair of dense GEMMs per tile on tcgen05 with the dq~ of one block riding along --, the merge / reduce
// kernels for the per-CTA partial states, and the host launchers.
#include "common.cuh"
#include "launchers.h"
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int kD = 256;
__device__ __forceinline__ void bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
// pointer arithmetic (not an integer round trip) so the compiler keeps the shared address space
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

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

size_t pool_fwd_workspace_bytes(int B, int max_len, int P) {
  int ns, tps;
  pool_tc_split_plan(max_len, B, &ns, &tps);
  const int PT = pool_tc_pad(P);
  return ((size_t)B * ns * PT * kD + (size_t)B * ns * 2 * PT) * sizeof(float);
}

int launch_pool_fwd(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* qt,
                    long long qt_stride, int P, float* workspace, float* pooled, float* lse, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  if (P <= 0 || P > 64) IMP_FAIL(IMP_ERR_ARG, "pool_fwd: P=%d out of [1,64]", P);
  if (total_rows <= 0 || max_len <= 0) IMP_FAIL(IMP_ERR_ARG, "pool_fwd: empty input (rows=%d, max_len=%d)", total_rows, max_len);
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

// dq~ partials first (B x max(nsplit_dq * PT, nsplit_dz * PP) x 256 floats), then the db1 partials of the dz pass
static size_t dq_part_elems(int B, int max_len, int P) {
  int nt, tt, nz, tz;
  pool_tc_split_plan(max_len, B, &nt, &tt);
  dz_split_plan(max_len, B, &nz, &tz);
  return std::max((size_t)B * nt * pool_tc_pad(P) * kD, (size_t)B * nz * pad_protos(P) * kD);
}

size_t pool_bwd_workspace_bytes(int B, int max_len, int P) {
  int nz, tz;
  dz_split_plan(max_len, B, &nz, &tz);
  return (dq_part_elems(B, max_len, P) + (size_t)B * nz * kD) * sizeof(float);
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
  float* part_dq = workspace;
  int rc, dq_ns, dq_pp;
  if (!dz) {
    // dq~ only (the later block of a stack): one streaming tcgen05 pass over h
    int nt, tt;
    pool_tc_split_plan(max_len, B, &nt, &tt);
    rc = launch_pool_tc_dq(h, total_rows, cu, B, qt[dq_block], qt_stride[dq_block], dpool[dq_block], (long long)P * kD,
                           lse[dq_block], delta[dq_block], P, nt, tt, part_dq, st);
    if (rc) return rc;
    dq_ns = nt; dq_pp = pool_tc_pad(P);
  } else {
    // dz (and db1) of all blocks at once; dq~ of `dq_block` rides along in the same pass
    DzParams z;
    dz_split_plan(max_len, B, &z.nsplit, &z.tiles_per_split);
    z.cu = cu; z.P = P; z.relu_mask = relu_mask; z.keep_scale = relu_mask ? keep_scale : 1.f; z.dz = dz;
    for (int k = 0; k < 2; ++k) {
      const int s = k < nblocks ? k : 0;
      z.qt[k] = qt[s]; z.qt_stride[k] = qt_stride[s]; z.dpool[k] = dpool[s]; z.dpool_stride[k] = (long long)P * kD;
      z.lse[k] = lse[s]; z.delta[k] = delta[s];
    }
    z.part_db = db1 ? workspace + dq_part_elems(B, max_len, P) : nullptr;
    z.part_dq = part_dq;
    z.dq_block = dq_block;
    CUtensorMap tmz;
    if ((rc = imp_make_tmap_2d(&tmz, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kZM,
                               CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if (PP == 16) rc = nblocks == 1 ? run_dz<16, 1>(tmz, z, B, st) : run_dz<16, 2>(tmz, z, B, st);
    else if (PP == 32) rc = nblocks == 1 ? run_dz<32, 1>(tmz, z, B, st) : run_dz<32, 2>(tmz, z, B, st);
    else rc = run_dz<64, 1>(tmz, z, B, st);
    if (rc) return rc;
    IMP_LAUNCH("zero_tail_rows", st, zero_tail_rows_kernel<<<imp_num_sms(), 256, 0, st>>>(dz, cu, B, total_rows));
    if (z.part_db) {
      IMP_LAUNCH("reduce_db", st, reduce_db_kernel<<<1, kD, 0, st>>>(z.part_db, db1, B * z.nsplit, db_accumulate));
    }
    dq_ns = z.nsplit; dq_pp = PP;
  }
  IMP_LAUNCH("reduce_dq", st, reduce_dq_kernel<<<dim3(P, B), kD, 0, st>>>(part_dq, dq, P, dq_pp, dq_ns));
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
