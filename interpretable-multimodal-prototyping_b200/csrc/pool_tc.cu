// A2/A3  softmax pooling of patch tokens into prototype tokens on tcgen05: forward and the dq~ pass.
// Reference: medmm/modeling/models/umeml_gan.py:65-80 and medmm/modeling/ops/attention.py:509-530
// (S = q k^T, softmax over the patches, A v) with the folded query of SURVEY.md 8 A3, so only h is streamed.
//
// One kernel, two modes, per 128-row tile of h (R,256) bf16 (TMA, 128-byte swizzle):
//   MMA1  S[n][c]      = sum_f h[n][f] G[c][f]          M 128 (patch rows), N = PP (fwd: q~) or 2 PP (dq: [q~ ; dpooled]),
//                                                        K 256; A = the h tile (K-major), B = G (K-major)
//   epilogue (8 warps: thread = patch row = TMEM lane x one half of the prototype columns)
//         fwd : w[n][p]  = exp2(S log2e - m_p)           m_p = the CTA's running reference maximum (see below)
//         dq  : w[n][p]  = a (dA - delta_p),  a = exp2(S log2e - lse_p log2e)
//         -> bf16, written as the MN-major B operand of MMA2 ([128 rows n][64 p], one 128-byte row per patch)
//   MMA2  acc[f][p]   += sum_n h[n][f] w[n][p]           M 128 (features, two halves), N 64, K 128 (patch rows);
//                                                        A = the SAME h tile read MN-major (h^T), B = w
//         fwd only:  l[p] += sum_n w[n][p]                from the same bf16-rounded weights: a 31-shuffle transposing
//                                                        reduction per warp leaves column p's sum in lane p
// The P x 256 results live transposed in TMEM (lane = feature, column = prototype): 2 x 64 columns instead of
// 256 x 128 lanes of which only P would be used, and a rescale touches 64 columns per thread.
//
// Running maximum without a per-tile rescale: m_p only moves when some score of the tile exceeds it by more than
// 8 (log2 units; weights then stay below 2^8, exact in bf16/fp32 terms).  Every tile the 128 epilogue threads vote
// (bar.red.or); on a hit they take the slow path: column maxima by warp shuffles, new m_p, and acc / l rescaled in
// TMEM (tcgen05.ld -> mul -> tcgen05.st) once the previous tile's MMA2 has retired.  The first tile always takes it
// (m = -inf); after that it is rare (the maximum over 128 patches of a bag is almost never 8 log2 units above the
// maximum over the bag's earlier tiles).  The partial state a CTA leaves is (m_p, l_p, acc_p) like the mma.sync
// kernel it replaces, merged across the CTAs of a bag by pool_merge / reduce_dq (pool.cu).
//
// Warps: 0-7 epilogue (with 4 the per-tile work of ~700 dependent instructions sat on one warp per scheduler and
// took ~2100 cycles; two warps per scheduler halve the columns per thread and hide each other's latencies),
// 8 TMA producer, 9 MMA issuer.  The issuer polls its barriers instead of waiting in a fixed
// order: MMA1 of tile t+1 goes out as soon as its h tile has landed (S is double buffered in TMEM), MMA2 of tile t as
// soon as w is written -- a blocking wait for the NEXT tile's load in front of MMA2(t) kept the stage of tile t, and with
// it the load of tile t+2, hostage (one load in flight per SM: 2.9 us per tile, first version).  The forward with
// P <= 32 runs three stages of h (192 KB).
#include "common.cuh"
#include "launchers.h"

namespace {

constexpr int kD = 256;
constexpr int kTM = 128;                        // patch rows per tile
constexpr int kTile = kTM * kD * 2;             // 64 KB: four [128 rows][64 feats] boxes
constexpr int kBox = kTM * 128;                 // 16 KB
constexpr int kEpiWarps = 8;                    // (TMEM lane quarter) x (column half)
constexpr int kTmaWarp = 8, kMmaWarp = 9;
constexpr int kThreads = 10 * 32;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kWBox = kTM * 128;                // one w buffer: [128 rows][64 p] bf16
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLazy = 8.f;                    // log2 units

enum { MODE_FWD = 0, MODE_DQ = 1 };

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// barrier over `n` threads that also ORs a predicate
__device__ __forceinline__ bool bar_or(int id, int n, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, %2, %3, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}"
      : "=r"(out)
      : "r"((uint32_t)pred), "r"(id), "r"(n)
      : "memory");
  return out != 0;
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct PoolTcParams {
  const int* cu;
  const float* qt;      long long qt_stride;     // (B|1, P, 256) folded queries
  const float* dpool;   long long dpool_stride;  // dq mode: (B, P, 256) token cotangents
  const float* lse;     // dq mode: (B, P)
  const float* delta;   // dq mode: (B, P)
  float* part_acc;      // (B, nsplit, PP, 256): fwd: sum_n w h (unnormalised) ; dq: partial dq~
  float* part_ml;       // fwd: (B, nsplit, 2, PP): m (natural log units), l
  int P, nsplit, tiles_per_split;
};

template <int PP, int MODE>
struct Cfg {
  static constexpr int N1 = MODE == MODE_FWD ? PP : 2 * PP;           // columns of S
  static constexpr int STAGES = (MODE == MODE_FWD && PP == 32) ? 3 : 2;   // h tiles in flight (shared memory budget)
  static constexpr int NBUF = (STAGES == 3 || (MODE == MODE_DQ && PP == 64)) ? 1 : 2;   // w buffers
  static constexpr size_t smem = 1024 + (size_t)STAGES * kTile + (size_t)N1 * 512 + (size_t)NBUF * kWBox + 6 * 64 * 4 + 192;
};

template <int PP, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
pool_tc_kernel(const __grid_constant__ CUtensorMap tm_h, const PoolTcParams p) {
  using C = Cfg<PP, MODE>;
  constexpr int N1 = C::N1, NBUF = C::NBUF, kStages = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* tiles = smem;                                       // kStages h tiles
  uint8_t* s_g = tiles + (size_t)kStages * kTile;              // G: 4 boxes [N1][64], box stride N1*128
  uint8_t* s_w = s_g + (size_t)N1 * 512;                       // NBUF boxes [128][64]
  float* s_wmax = reinterpret_cast<float*>(s_w + (size_t)NBUF * kWBox);   // [4][64]
  float* s_m = s_wmax + 4 * 64;                                // [64] running maximum (log2 units) | dq: lse * log2e
  float* s_alpha = s_m + 64;                                   // [64] rescale factors            | dq: delta
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_alpha + 64);
  uint64_t* full = bars + 14;       // [kStages] TMA -> MMA1
  uint64_t* empty = bars + 17;      // [kStages] MMA2 retired -> TMA
  uint64_t* sfull = bars + 4;       // [2] MMA1 -> epilogue
  uint64_t* sempty = bars + 6;      // [2] 8 warps -> MMA1 (S in registers)
  uint64_t* wfull = bars + 8;       // [2] 8 warps -> MMA2 (w written, acc rescaled)
  uint64_t* wempty = bars + 10;     // [2] MMA2 retired -> epilogue (w buffer reusable)
  uint64_t* accdone = bars + 12;    // MMA2 of a tile retired (slow path: at most one phase behind, see there)
  uint64_t* alldone = bars + 13;    // MMA2 of the CTA's LAST tile retired (single phase)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int b = blockIdx.y, split = blockIdx.x;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int ntiles_bag = (row_end - row_begin + kTM - 1) / kTM;
  const int t0 = split * p.tiles_per_split;
  const int ntiles = max(0, min(ntiles_bag, t0 + p.tiles_per_split) - t0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* out_acc = p.part_acc + ((size_t)b * p.nsplit + split) * PP * kD;
  if (ntiles == 0) {                               // neutral partial state
    if (MODE == MODE_FWD) {
      float* out_ml = p.part_ml + ((size_t)b * p.nsplit + split) * 2 * PP;
      if (threadIdx.x < PP) { out_ml[threadIdx.x] = -INFINITY; out_ml[PP + threadIdx.x] = 0.f; }
    } else {
      for (int i = threadIdx.x; i < PP * kD; i += kThreads) out_acc[i] = 0.f;
    }
    return;
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tm_h);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sfull[i], 1); mbar_init(&sempty[i], kEpiWarps); mbar_init(&wfull[i], kEpiWarps); mbar_init(&wempty[i], 1); }
    mbar_init(accdone, 1);
    mbar_init(alldone, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();                               // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s0 = tmem_base, tm_acc = tmem_base + 256;   // S: 2 x N1 <= 256 columns; acc: 2 x 64
  if (warp != kTmaWarp) {
    // The h tiles are already on their way (the TMA warp below does not wait for this): stage G and the small tables
    // with the other nine warps.  Loads are issued four deep -- a load -> pack -> store chain per item made this
    // prologue 10 % of the kernel (one L2 round trip per iteration).
    const int tid = warp == kMmaWarp ? kEpiThreads + lane : threadIdx.x;       // 288 workers
    constexpr int kWorkers = kEpiThreads + 32, kItems = N1 * 64;
    for (int base = tid; base < kItems; base += 4 * kWorkers) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * kWorkers;
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < kItems) {
          const int r = i >> 6, c = (i & 63) << 2, pi = r % PP;
          if (pi < p.P) {
            const float* src = r >= PP ? p.dpool + (size_t)b * p.dpool_stride : p.qt + (size_t)b * p.qt_stride;
            v[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)pi * kD + c));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * kWorkers;
        if (i < kItems) {
          const int r = i >> 6, c = (i & 63) << 2;
          const uint32_t off = (uint32_t)((c >> 6) * (N1 * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4) + ((c & 7) << 1));
          *reinterpret_cast<uint2*>(s_g + off) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
        }
      }
    }
    for (int i = tid; i < NBUF * kWBox / 16; i += kWorkers) reinterpret_cast<uint4*>(s_w)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (MODE == MODE_FWD) {
      for (int i = tid; i < 64; i += kWorkers) { s_m[i] = -INFINITY; s_alpha[i] = 0.f; }
    } else {
      for (int i = tid; i < 64; i += kWorkers) {
        const bool ok = i < p.P;
        s_m[i] = ok ? p.lse[(size_t)b * p.P + i] * kLog2e : INFINITY;       // padded prototypes: a = 0
        s_alpha[i] = ok ? p.delta[(size_t)b * p.P + i] : 0.f;
      }
    }
    fence_proxy_async_smem();                    // G / zeroed w: generic-proxy writes read by the MMAs
    bar_sync(3, kWorkers);
  }

  if (warp == kTmaWarp) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int i = 0; i < ntiles; ++i) {
        const int stage = i % kStages;
        mbar_wait_idle(&empty[stage], ((i / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kTile);
        uint8_t* dst = tiles + (size_t)stage * kTile;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx)
          tma_load_2d(dst + bx * kBox, &tm_h, &full[stage], bx * 64, row_begin + (t0 + i) * kTM);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kTM, N1, 0, 0);
      constexpr uint32_t idesc2 = umma_idesc_bf16(128, 64, 1, 1);       // A = h^T (MN-major), B = w (MN-major)
      const uint32_t sg = smem_u32(s_g), sw = smem_u32(s_w);
      int next1 = 0, next2 = 0;                                           // next tile for MMA1 / MMA2
      while (next2 < ntiles) {
        // MMA1(j): its h tile has landed and the S buffer was read.  Never more than one tile ahead of the MMA2 that
        // was ISSUED last (the slow path's accdone parity argument needs MMA2(j-2) in front of MMA1(j)).
        if (next1 < ntiles && next1 <= next2 + 1 && mbar_test(&full[next1 % kStages], (next1 / kStages) & 1) &&
            mbar_test(&sempty[next1 & 1], ((next1 >> 1) & 1) ^ 1)) {
          const int j = next1++, stage = j % kStages, buf = j & 1;
          tc_fence_after();
          const uint32_t sh = smem_u32(tiles + (size_t)stage * kTile);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint64_t ad = umma_desc_sw128(sh + (k >> 2) * kBox + (k & 3) * 32, 0, 1024);
            const uint64_t bd = umma_desc_sw128(sg + (k >> 2) * (N1 * 128) + (k & 3) * 32, 0, 1024);
            umma_f16(tm_s0 + buf * N1, ad, bd, idesc1, k != 0);
          }
          umma_commit(&sfull[buf]);
          continue;
        }
        if (next2 < next1 && mbar_test(&wfull[next2 % NBUF], (next2 / NBUF) & 1)) {
          const int i = next2++, stage = i % kStages, wb = i % NBUF;
          tc_fence_after();
          const uint32_t sh = smem_u32(tiles + (size_t)stage * kTile);
          const uint32_t swb = sw + wb * kWBox;
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
            for (int k = 0; k < kTM / 16; ++k) {
              // h^T: two 64-feature boxes (16 KB apart) per 128-feature half, 16 patch rows = 2048 B per k-step
              const uint64_t ad = umma_desc_sw128(sh + hf * 2 * kBox + k * 2048, kBox, 1024);
              const uint64_t bd = umma_desc_sw128(swb + k * 2048, kWBox, 1024);
              umma_f16(tm_acc + hf * 64, ad, bd, idesc2, (i | k) != 0);
            }
          }
          umma_commit(&empty[stage]);
          umma_commit(&wempty[wb]);
          umma_commit(accdone);
          if (i == ntiles - 1) umma_commit(alldone);
          continue;
        }
        __nanosleep(32);
      }
    }
  } else {
    // ------------------------------ epilogue warps ------------------------------
    constexpr int CH = PP / 2;                            // prototype columns per thread
    const int q = warp & 3, hf = warp >> 2;
    const int n = q * 32 + lane;                          // row inside the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int col0 = hf * CH;
    const int mycol = col0 + (lane & (CH - 1));           // fwd: the column whose partial normaliser this lane keeps
    uint8_t* wrow_base = s_w + n * 128;
    float lpart = 0.f;
    auto ld_cols = [&](uint32_t addr, uint32_t (&v)[CH]) {
      if constexpr (CH == 16) tmem_ld16(addr, v); else tmem_ld32(addr, v);
    };
    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1, wb = i % NBUF;
      const bool row_ok = row_begin + (t0 + i) * kTM + n < row_end;
      mbar_wait(&sfull[buf], (i >> 1) & 1);
      tc_fence_after();
      float w[CH];
      if (MODE == MODE_FWD) {
        float t[CH];
        {
          uint32_t v[CH];
          ld_cols(tm_s0 + buf * N1 + lane_addr + col0, v);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < CH; ++e) t[e] = row_ok ? __uint_as_float(v[e]) * kLog2e : -INFINITY;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[buf]);
        bool hit = false;
#pragma unroll
        for (int e = 0; e < CH; ++e) hit |= t[e] > s_m[col0 + e] + kLazy;
        if (bar_or(1, kEpiThreads, hit)) {
          // ---- slow path: new reference maxima, rescale of the accumulators ----
#pragma unroll
          for (int e = 0; e < CH; ++e) {
            const float mx = warp_max(t[e]);
            if (lane == (e & 31)) s_wmax[q * 64 + col0 + e] = mx;
          }
          bar_sync(2, kEpiThreads);
          if (threadIdx.x < PP) {
            const int c = threadIdx.x;
            const float mx = fmaxf(fmaxf(s_wmax[c], s_wmax[64 + c]), fmaxf(s_wmax[128 + c], s_wmax[192 + c]));
            const float m_old = s_m[c];
            const float m_new = fmaxf(m_old, mx);                       // finite: the tile has a valid row
            s_alpha[c] = exp2f(m_old - m_new);                          // 0 on the first tile (m_old = -inf)
            s_m[c] = m_new;
          }
          bar_sync(2, kEpiThreads);
          lpart *= s_alpha[mycol];
          if (i > 0) {
            // MMA2 of tile i-1 retired.  S of tile i exists, so MMA1(i) and everything issued before it, MMA2(i-2)
            // included, has completed (the tensor pipe retires in issue order): accdone has seen i-1 or i commits and
            // the parity of phase i-1 is unambiguous.
            mbar_wait(accdone, (i - 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < PP; c0 += 32) {                       // this warp's quarter of acc half `hf`
              uint32_t v[32];
              const uint32_t addr = tm_acc + hf * 64 + lane_addr + c0;
              tmem_ld32(addr, v);
              tmem_ld_wait();
#pragma unroll
              for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * s_alpha[c0 + e]);
              tmem_st32(addr, v);
            }
            tmem_st_wait();
            tc_fence_before();
          }
        }
#pragma unroll
        for (int e = 0; e < CH; ++e) w[e] = exp2f(t[e] - s_m[col0 + e]);
      } else {
        uint32_t sv[CH], av[CH];
        ld_cols(tm_s0 + buf * N1 + lane_addr + col0, sv);
        ld_cols(tm_s0 + buf * N1 + lane_addr + PP + col0, av);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < CH; ++e) {
          const float a = row_ok ? exp2f(__uint_as_float(sv[e]) * kLog2e - s_m[col0 + e]) : 0.f;
          w[e] = a * (__uint_as_float(av[e]) - s_alpha[col0 + e]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[buf]);
      }
      uint32_t pk[CH / 2];
#pragma unroll
      for (int e = 0; e < CH / 2; ++e) pk[e] = pack_bf16x2(w[2 * e], w[2 * e + 1]);
      if (MODE == MODE_FWD) {
        // normaliser from the ROUNDED weights: transposing butterfly over the warp's 32 rows; lane l ends with the sum of
        // column (l mod CH) (CH = 16: 15 shuffles + one across the two half-warps; CH = 32: 31 shuffles)
        float v[CH];
#pragma unroll
        for (int e = 0; e < CH / 2; ++e) { v[2 * e] = bf16lo(pk[e]); v[2 * e + 1] = bf16hi(pk[e]); }
#pragma unroll
        for (int sft = CH / 2; sft >= 1; sft >>= 1) {
          const bool up = (lane & sft) != 0;
#pragma unroll
          for (int j = 0; j < sft; ++j) {
            const float send = up ? v[j] : v[j + sft];
            const float keep = up ? v[j + sft] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
          }
        }
        if (CH == 16) v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
        lpart += v[0];
      }
      // ---- w -> shared memory (MN-major B operand of MMA2), once the MMA2 that last read this buffer retired ----
      mbar_wait(&wempty[wb], ((i / NBUF) & 1) ^ 1);
      uint8_t* wrow = wrow_base + wb * kWBox;
#pragma unroll
      for (int c = 0; c < CH / 8; ++c)
        *reinterpret_cast<uint4*>(wrow + ((((col0 >> 3) + c) ^ (n & 7)) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&wfull[wb]);
    }
    // ---- write the partial state: acc^T lives as [feature lane][prototype column] ----
    // the epilogue can run two MMA2 commits ahead here, which a parity wait on accdone cannot tell apart
    mbar_wait(alldone, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < PP; c0 += 32) {                                 // feature half hf, features q*32 + lane
      uint32_t v[32];
      tmem_ld32(tm_acc + hf * 64 + lane_addr + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) out_acc[(size_t)(c0 + e) * kD + hf * 128 + n] = __uint_as_float(v[e]);
    }
    if (MODE == MODE_FWD) {
      float* out_ml = p.part_ml + ((size_t)b * p.nsplit + split) * 2 * PP;
      if (lane < CH) s_wmax[q * 64 + mycol] = lpart;
      bar_sync(2, kEpiThreads);
      if (threadIdx.x < PP) {
        const int c = threadIdx.x;
        out_ml[c] = s_m[c] * kLn2;
        out_ml[PP + c] = s_wmax[c] + s_wmax[64 + c] + s_wmax[128 + c] + s_wmax[192 + c];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int PP, int MODE>
int run_pool_tc(const CUtensorMap& tm, const PoolTcParams& p, int B, cudaStream_t st) {
  constexpr size_t smem = Cfg<PP, MODE>::smem;
  static_assert(smem <= 227 * 1024, "pool_tc shared memory");
  { const int rc_ = imp_ensure_smem((const void*)pool_tc_kernel<PP, MODE>, smem); if (rc_) return rc_; }
  if (MODE == MODE_FWD) {
    IMP_LAUNCH("pool_fwd", st, pool_tc_kernel<PP, MODE><<<dim3(p.nsplit, B), kThreads, smem, st>>>(tm, p));
  } else {
    IMP_LAUNCH("pool_bwd_dq", st, pool_tc_kernel<PP, MODE><<<dim3(p.nsplit, B), kThreads, smem, st>>>(tm, p));
  }
  return IMP_OK;
}

}  // namespace

int pool_tc_pad(int P) { return P <= 32 ? 32 : 64; }

// 128-row tiles, one CTA per SM, at least 4 tiles (256 KB of h) per CTA; whole waves (pool.cu best_split)
void best_split(int tiles, int B, int slots, int min_tiles, int max_split, int* nsplit, int* tiles_per_split);
void pool_tc_split_plan(int max_len, int B, int* nsplit, int* tiles_per_split) {
  const int tiles = max(1, (max_len + kTM - 1) / kTM);
  best_split(tiles, B, imp_num_sms(), 4, 128, nsplit, tiles_per_split);
}

// forward partials: part_acc (B, nsplit, PP, 256), part_ml (B, nsplit, 2, PP)
int launch_pool_tc_fwd(const bf16* h, int total_rows, const int* cu, int B, const float* qt, long long qt_stride, int P,
                       int nsplit, int tiles_per_split, float* part_acc, float* part_ml, cudaStream_t st) {
  PoolTcParams p;
  p.cu = cu; p.qt = qt; p.qt_stride = qt_stride; p.dpool = nullptr; p.dpool_stride = 0; p.lse = nullptr; p.delta = nullptr;
  p.part_acc = part_acc; p.part_ml = part_ml; p.P = P; p.nsplit = nsplit; p.tiles_per_split = tiles_per_split;
  CUtensorMap tm;
  int rc = imp_make_tmap_2d(&tm, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kTM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  return pool_tc_pad(P) == 32 ? run_pool_tc<32, MODE_FWD>(tm, p, B, st) : run_pool_tc<64, MODE_FWD>(tm, p, B, st);
}

// dq~ partials of one pooling block: part_dq (B, nsplit, PP, 256)
int launch_pool_tc_dq(const bf16* h, int total_rows, const int* cu, int B, const float* qt, long long qt_stride,
                      const float* dpool, long long dpool_stride, const float* lse, const float* delta, int P, int nsplit,
                      int tiles_per_split, float* part_dq, cudaStream_t st) {
  PoolTcParams p;
  p.cu = cu; p.qt = qt; p.qt_stride = qt_stride; p.dpool = dpool; p.dpool_stride = dpool_stride; p.lse = lse; p.delta = delta;
  p.part_acc = part_dq; p.part_ml = nullptr; p.P = P; p.nsplit = nsplit; p.tiles_per_split = tiles_per_split;
  CUtensorMap tm;
  int rc = imp_make_tmap_2d(&tm, h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, total_rows, kD * 2, 64, kTM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  return pool_tc_pad(P) == 32 ? run_pool_tc<32, MODE_DQ>(tm, p, B, st) : run_pool_tc<64, MODE_DQ>(tm, p, B, st);
}
