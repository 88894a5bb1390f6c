// A4-A6  modularity loss of a token set against the patch graph of a bag, forward and the
// gradient wrt the (normalised) tokens in one sweep.
// Reference: medmm/modeling/ops/utils.py:178-228 (cluster_assignment_matrix,
// get_modularity_matrix_and_edge, compute_modularity), called at umeml_gan.py:516-529.
//
// Per bag (N patches, x = h detached):   xh_i = x_i/|x_i| ,  A_ij = relu(xh_i.xh_j) [i != j]
//   d_i = sum_j A_ij , e = sum_i d_i ,  C_ip = relu(xh_i . chat_p)
//   delta_ij = tanh(max_p C_ip C_jp / temp)
//   loss = -100 [ sum_ij A_ij delta_ij / e  -  sum_ij d_i d_j delta_ij / e^2 ]      (== utils.py:220-228)
// The reference materialises P x N x N and an N^3 matmul; here nothing larger than N x P is stored.
//
// Kernels:
//   prep     : inv-norms, xh (bf16, the Gram operand) and L = fixed-point log2(C) per (patch, token)
//   gram<0>  : degrees d, e       (tcgen05 Gram tiles 128 x 64, K = 256, accumulators in TMEM)
//   gram<1>  : the two traces and T_ip = sum_j 2 g_ij (1-delta^2)/temp u_ij [p = argmax]
//              The (max,x) contraction over tokens runs on the integer pipe in the log domain:
//              max_p (L_ip + L_jp) with the token index carried in the 5 low bits (one
//              add-max instruction per (pair, token)), u = exp2(max).
//   finish   : dC = T / C, dchat_p = sum_i dC_ip xh_i, loss per token group
// Up to two token groups (<= 32 tokens each) share A, d and e (the reference evaluates the
// prototype tokens and the omic tokens against the same bag, umeml_gan.py:520-521).
#include "common.cuh"
#include "launchers.h"
#include <algorithm>
#include <stdlib.h>

namespace {

constexpr int kD = 256;
constexpr int kBM = 128;                 // rows (i) per CTA
constexpr int kBN = 64;                  // columns (j) per tile
constexpr int kABytes = kBM * kD * 2;    // 64 KB, four [128][64] swizzled boxes
constexpr int kBBytes = kBN * kD * 2;    // 32 KB, four [64][64] boxes
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kStages = 2;
constexpr int kLogShift = 16;            // fixed point: log2(C) * 2^16, then << 5 with the token index below
constexpr int kZeroFix = -30000000;      // "C == 0" marker (two of them still fit in int32 after << 5)

// pointer arithmetic (not an integer round trip) so the compiler keeps the shared address space
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------
// prep: one warp per patch row
// ------------------------------------------------------------------------------------------
struct PrepParams {
  const bf16* h;            // (R,256)
  const int* cu;
  const float* chat;        // (B, Pt, 256): tokens normalised across tokens per feature
  bf16* xh;                 // (Rpad,256)
  float* invn;              // (Rpad)
  int* lfix;                // (Rpad, PtPad)
  int B, P1, P2, P1pad, PtPad;
};

__global__ void __launch_bounds__(256) modularity_prep_kernel(const PrepParams p) {
  extern __shared__ float s_c[];                 // (PtPad, 256) fp32, zero rows for padding
  const int b = blockIdx.y;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int Pt = p.P1 + p.P2;
  for (int i = threadIdx.x; i < p.PtPad * kD; i += blockDim.x) {
    const int slot = i / kD, f = i % kD;
    int src = -1;
    if (slot < p.P1) src = slot;
    else if (slot >= p.P1pad && slot - p.P1pad < p.P2) src = p.P1 + slot - p.P1pad;
    s_c[i] = src >= 0 ? p.chat[((size_t)b * Pt + src) * kD + f] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = row_begin + blockIdx.x * 8 + warp; row < row_end; row += gridDim.x * 8) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p.h + (size_t)row * kD + lane * 8);
    float v[8] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y),
                  bf16lo(raw.z), bf16hi(raw.z), bf16lo(raw.w), bf16hi(raw.w)};
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) ss += v[k] * v[k];
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);          // F.normalize eps (utils.py:179,193)
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] *= inv;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p.xh + (size_t)row * kD + lane * 8) = o;
    if (lane == 0) p.invn[row] = inv;
    for (int slot = 0; slot < p.PtPad; ++slot) {
      const float4 c0 = *reinterpret_cast<const float4*>(s_c + slot * kD + lane * 8);
      const float4 c1 = *reinterpret_cast<const float4*>(s_c + slot * kD + lane * 8 + 4);
      float dot = v[0] * c0.x + v[1] * c0.y + v[2] * c0.z + v[3] * c0.w + v[4] * c1.x + v[5] * c1.y + v[6] * c1.z + v[7] * c1.w;
      dot = warp_sum(dot);
      if (lane == 0) {
        const bool real = slot < p.P1 || (slot >= p.P1pad && slot - p.P1pad < p.P2);
        int fix = kZeroFix;
        if (real && dot > 0.f) fix = max(kZeroFix, __float2int_rn(log2f(dot) * (float)(1 << kLogShift)));
        // low 5 bits: token index inside its group (column-side operand); the row side masks them off
        p.lfix[(size_t)row * p.PtPad + slot] = fix * 32 + (slot < p.P1pad ? slot : slot - p.P1pad);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Gram sweep.  MODE 0: degrees.  MODE 1: traces + T.   NQ1/NQ2: int4 quads per token group.
// ------------------------------------------------------------------------------------------
struct GramParams {
  const int* cu;
  const int* lfix;          // (Rpad, PtPad)
  float* d;                 // (Rpad) degrees: written (atomicAdd) in MODE 0, read in MODE 1
  double* e;                // (B)
  float* T;                 // (Rpad, PtPad) atomicAdd
  double* s;                // (B, 2 groups, 2): s1, s2
  int tiles_per_split;      // column tiles per CTA
  float inv_temp;
};

template <int MODE, int NQ1, int NQ2>
constexpr size_t gram_smem() {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  size_t stage = kBBytes + (MODE ? (size_t)kBN * PtPad * 4 : 0);
  size_t t = MODE ? (size_t)kEpiWarps * 32 * (PtPad + 1) * 4 : 0;
  return 1024 + kABytes + kStages * ((stage + 1023) & ~(size_t)1023) + t + 256;
}

template <int MODE, int NQ1, int NQ2>
__global__ void __launch_bounds__(kThreads, 1)
modularity_gram_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                       const GramParams p) {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  constexpr int kLBytes = MODE ? kBN * PtPad * 4 : 0;
  constexpr int kStageBytes = (kBBytes + kLBytes + 1023) & ~1023;
  constexpr int TS = PtPad + 1;                       // row stride of the private T accumulators
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_a = smem;
  uint8_t* s_stage = s_a + kABytes;
  float* s_T = reinterpret_cast<float*>(s_stage + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_T) + (MODE ? kEpiWarps * 32 * TS * 4 : 0));
  uint64_t* full = bars;                  // [kStages]  TMA -> MMA + epilogue
  uint64_t* empty = bars + kStages;       // [kStages]  MMA commit + 8 epilogue warps -> TMA
  uint64_t* tfull = bars + 2 * kStages;   // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;           // [2] epilogue -> MMA
  uint64_t* afull = tempty + 2;           // A block landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);   // [8][4] block reduction scratch

  const int b = blockIdx.z;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int n = row_end - row_begin;
  const int i0 = row_begin + blockIdx.x * kBM;
  if (i0 >= row_end) return;
  const int ntiles_bag = (n + kBN - 1) / kBN;
  const int t0 = blockIdx.y * p.tiles_per_split;
  const int ntiles = max(0, min(ntiles_bag, t0 + p.tiles_per_split) - t0);
  if (ntiles == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MODE ? 1 + kEpiWarps : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    mbar_init(afull, 1);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiWarps) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(afull, kABytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_a + bx * (kBM * 128), &tm_a, afull, bx * 64, i0);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % kStages;
        mbar_wait_idle(&empty[stage], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kBBytes + kLBytes);
        uint8_t* dst = s_stage + (size_t)stage * kStageBytes;
        const int j0 = row_begin + (t0 + it) * kBN;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tma_load_2d(dst + bx * (kBN * 128), &tm_b, &full[stage], bx * 64, j0);
        if (MODE) bulk_load(dst + kBBytes, p.lfix + (size_t)j0 * PtPad, kLBytes, &full[stage]);
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
      mbar_wait_idle(afull, 0);
      tc_fence_after();
      const uint32_t sa = smem_u32(s_a);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % kStages, acc = it & 1;
        mbar_wait_idle(&tempty[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait_idle(&full[stage], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t sb = smem_u32(s_stage + (size_t)stage * kStageBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(sa + (k >> 2) * (kBM * 128) + (k & 3) * 32, 0, 1024);
          const uint64_t bd = umma_desc_sw128(sb + (k >> 2) * (kBN * 128) + (k & 3) * 32, 0, 1024);
          umma_f16(tmem_base + acc * kBN, ad, bd, idesc, k != 0);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------ epilogue: thread = (row i, 32 of the 64 columns) ------------------
    const int q = warp & 3, hc = warp >> 2;
    const int i = i0 + q * 32 + lane;
    const bool row_ok = i < row_end;
    float acc_d = 0.f;                               // MODE 0
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};    // MODE 1
    int Li[MODE ? PtPad : 1];
    float di = 0.f, k1 = 0.f, k2 = 0.f;
    float* myT = s_T + (size_t)(warp * 32 + lane) * TS;
    if (MODE) {
#pragma unroll
      for (int k = 0; k < PtPad; ++k) {
        Li[k] = row_ok ? (__ldg(p.lfix + (size_t)i * PtPad + k) & ~31) : (kZeroFix * 32);
        myT[k] = 0.f;
      }
      di = row_ok ? __ldg(p.d + i) : 0.f;
      const double e = p.e[b];
      k1 = (float)(1.0 / e);
      k2 = (float)(1.0 / (e * e));
    }
    const float nx2scale = -2.f * 1.4426950408889634f * p.inv_temp;   // tanh(u/temp) via exp2(-2 u log2e / temp)
    const float gscale = -200.f * p.inv_temp;                        // 2 * (-100) / temp
    for (int it = 0; it < ntiles; ++it) {
      const int stage = it % kStages, acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + hc * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);      // accumulator is in registers: release TMEM early
      const int jbase = row_begin + (t0 + it) * kBN + hc * 32;
      if (MODE == 0) {
        // interior tiles (no diagonal, no bag end) need no per-pair masks: relu + add only
        const int jt0 = row_begin + (t0 + it) * kBN;
        const bool interior = (jt0 + kBN <= row_end) && (jt0 >= i0 + kBM || jt0 + kBN <= i0);
        if (interior) {
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int jj = 0; jj < 32; jj += 2) {
            p0 += fmaxf(__uint_as_float(v[jj]), 0.f);
            p1 += fmaxf(__uint_as_float(v[jj + 1]), 0.f);
          }
          acc_d += p0 + p1;
        } else {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) {
            const int j = jbase + jj;
            const float a = fmaxf(__uint_as_float(v[jj]), 0.f);
            acc_d += (j < row_end && j != i) ? a : 0.f;
          }
        }
      } else {
        mbar_wait(&full[stage], (it / kStages) & 1);           // L tile of this stage
        const uint8_t* st = s_stage + (size_t)stage * kStageBytes + kBBytes;
        const int4* sL = reinterpret_cast<const int4*>(st) + (size_t)hc * 32 * (PtPad / 4);
        const float dj_lane = (jbase + lane < row_end) ? __ldg(p.d + jbase + lane) : 0.f;
        // Two columns per step and four independent add-max chains per token group (a dependent
        // VIADDMNMX issues only every ~5 cycles), software-pipelined: the chains of columns jj+2,jj+3
        // are issued before the exp2/rcp tail of columns jj,jj+1 so the MUFU latency hides behind them.
        auto chains = [&](int jj, int (&mg0)[2], int (&mg1)[2]) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int4* lj = sL + (jj + c) * (PtPad / 4);
            int ch[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
            for (int qd = 0; qd < NQ1; ++qd) {
              const int4 w = lj[qd];
              ch[0] = __viaddmax_s32(Li[4 * qd + 0], w.x, ch[0]);
              ch[1] = __viaddmax_s32(Li[4 * qd + 1], w.y, ch[1]);
              ch[2] = __viaddmax_s32(Li[4 * qd + 2], w.z, ch[2]);
              ch[3] = __viaddmax_s32(Li[4 * qd + 3], w.w, ch[3]);
            }
            mg0[c] = max(__vimax3_s32(ch[0], ch[1], ch[2]), ch[3]);
            int cg[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
            for (int qd = 0; qd < NQ2; ++qd) {
              const int4 w = lj[NQ1 + qd];
              cg[0] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 0], w.x, cg[0]);
              cg[1] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 1], w.y, cg[1]);
              cg[2] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 2], w.z, cg[2]);
              cg[3] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 3], w.w, cg[3]);
            }
            mg1[c] = max(__vimax3_s32(cg[0], cg[1], cg[2]), cg[3]);
          }
        };
        auto tail = [&](int jj, const int (&mg0)[2], const int (&mg1)[2]) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int j = jbase + jj + c;
            const bool ok = row_ok && j < row_end;
            const float a = (ok && j != i) ? fmaxf(__uint_as_float(v[jj + c]), 0.f) : 0.f;
            const float djv = __shfl_sync(0xffffffffu, dj_lane, jj + c);   // warp-uniform: never under a lane predicate
            const float dd = ok ? di * djv : 0.f;
            const float gw4 = 4.f * gscale * (a * k1 - dd * k2);           // 0 for masked pairs
#pragma unroll
            for (int grp = 0; grp < (NQ2 ? 2 : 1); ++grp) {
              const int m = grp ? mg1[c] : mg0[c];
              const int pstar = (m & 31) + (grp ? 4 * NQ1 : 0);
              // the 5 index bits perturb log2(u) by < 2^-16: far below the fixed-point resolution that matters
              const float u = ex2_approx((float)m * (1.f / (float)(1 << (kLogShift + 5))));
              const float e2 = ex2_approx(u * nx2scale);             // exp(-2u/temp)
              const float r = rcp_approx(1.f + e2);
              const float t1 = e2 * r;
              const float delta = r - t1;                            // tanh(u/temp) = (1-e2)/(1+e2)
              s1[grp] = fmaf(a, delta, s1[grp]);
              s2[grp] = fmaf(dd, delta, s2[grp]);
              myT[pstar] += (gw4 * t1) * (r * u);                    // 2 g (1-delta^2)/temp * u,  1-delta^2 = 4 e2 r^2
            }
          }
        };
        int ma0[2], ma1[2];
        chains(0, ma0, ma1);
#pragma unroll
        for (int jj = 0; jj < 32; jj += 2) {
          int mb0[2] = {INT_MIN, INT_MIN}, mb1[2] = {INT_MIN, INT_MIN};
          if (jj + 2 < 32) chains(jj + 2, mb0, mb1);
          tail(jj, ma0, ma1);
          ma0[0] = mb0[0]; ma0[1] = mb0[1]; ma1[0] = mb1[0]; ma1[1] = mb1[1];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
    }
    // ---------------- flush ----------------
    if (MODE == 0) {
      if (row_ok) atomicAdd(p.d + i, acc_d);
      float tot = warp_sum(row_ok ? acc_d : 0.f);
      if (lane == 0) s_red[warp] = tot;
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kEpiWarps; ++w) t += (double)s_red[w];
        atomicAdd(p.e + b, t);
      }
    } else {
      if (row_ok) {
#pragma unroll 4
        for (int k = 0; k < PtPad; ++k) {
          const float t = myT[k];
          if (t != 0.f) atomicAdd(p.T + (size_t)i * PtPad + k, t);
        }
      }
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        const float a1 = warp_sum(s1[grp]), a2 = warp_sum(s2[grp]);
        if (lane == 0) { s_red[warp * 4 + grp * 2] = a1; s_red[warp * 4 + grp * 2 + 1] = a2; }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < kEpiWarps; ++w) t += (double)s_red[w * 4 + threadIdx.x];
        atomicAdd(p.s + (size_t)b * 4 + threadIdx.x, t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------
// Symmetric pair sweep (traces + T).  Every unordered patch pair {i,j}, i <= j, is evaluated once by
// the CTA that owns row i: CTA = 128 rows x the 32-column tiles at or right of its diagonal.
//   thread = (row i, 16 of the 32 columns).  Row side: T_i[p*] += t in a private smem row.
//   Column side: (t | p*) words go through a warp-private smem transpose so that lane = column owns
//   the update T_j[p*] += t (no atomics in the loop); column accumulators are flushed to global
//   memory per tile with one RED per non-zero entry.
// ------------------------------------------------------------------------------------------
constexpr int kPN = 32;                               // columns per tile of the pair sweep
constexpr int kPBBytes = kPN * kD * 2;                // 16 KB, four [32][64] boxes
constexpr int kXWords = 2 * 16 * 33;                  // per warp: [group][column][row + pad]

template <int NQ1, int NQ2>
constexpr size_t pairs_smem() {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  size_t stage = (kPBBytes + (size_t)kPN * PtPad * 4 + 1023) & ~(size_t)1023;
  size_t t = 2 * (size_t)kEpiWarps * 32 * (PtPad + 1) * 4;
  return 1024 + kABytes + kStages * stage + t + (size_t)kEpiWarps * kXWords * 4 + 256;
}

template <int NQ1, int NQ2>
__global__ void __launch_bounds__(kThreads, 1)
modularity_pairs_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                        const GramParams p) {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  constexpr int kLBytes = kPN * PtPad * 4;
  constexpr int kStageBytes = (kPBBytes + kLBytes + 1023) & ~1023;
  constexpr int TS = PtPad + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_a = smem;
  uint8_t* s_stage = s_a + kABytes;
  float* s_Trow = reinterpret_cast<float*>(s_stage + kStages * kStageBytes);
  float* s_Tcol = s_Trow + kEpiWarps * 32 * TS;
  uint32_t* s_X = reinterpret_cast<uint32_t*>(s_Tcol + kEpiWarps * 32 * TS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_X + kEpiWarps * kXWords);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = tfull + 2;
  uint64_t* afull = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);

  const int b = blockIdx.x, rb = blockIdx.z;          // bags vary fastest: the heavy row blocks (rb = 0) start first
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int n = row_end - row_begin;
  const int i0 = row_begin + rb * kBM;
  if (i0 >= row_end) return;
  const int ntiles_bag = (n + kPN - 1) / kPN;
  const int t_lo = rb * (kBM / kPN);                       // first tile that touches the diagonal block
  const int per = (ntiles_bag - t_lo + (int)gridDim.y - 1) / (int)gridDim.y;
  const int t0 = t_lo + blockIdx.y * per;
  const int ntiles = max(0, min(ntiles_bag, t0 + per) - t0);
  if (ntiles == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + kEpiWarps); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    mbar_init(afull, 1);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiWarps) {
    if (lane == 0) {
      mbar_arrive_expect_tx(afull, kABytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_a + bx * (kBM * 128), &tm_a, afull, bx * 64, i0);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % kStages;
        mbar_wait_idle(&empty[stage], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kPBBytes + kLBytes);
        uint8_t* dst = s_stage + (size_t)stage * kStageBytes;
        const int j0 = row_begin + (t0 + it) * kPN;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tma_load_2d(dst + bx * (kPN * 128), &tm_b, &full[stage], bx * 64, j0);
        bulk_load(dst + kPBBytes, p.lfix + (size_t)j0 * PtPad, kLBytes, &full[stage]);
      }
    }
  } else if (warp == kEpiWarps + 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kPN, 0, 0);
      mbar_wait_idle(afull, 0);
      tc_fence_after();
      const uint32_t sa = smem_u32(s_a);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % kStages, acc = it & 1;
        mbar_wait_idle(&tempty[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait_idle(&full[stage], (it / kStages) & 1);
        tc_fence_after();
        const uint32_t sb = smem_u32(s_stage + (size_t)stage * kStageBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(sa + (k >> 2) * (kBM * 128) + (k & 3) * 32, 0, 1024);
          const uint64_t bd = umma_desc_sw128(sb + (k >> 2) * (kPN * 128) + (k & 3) * 32, 0, 1024);
          umma_f16(tmem_base + acc * kPN, ad, bd, idesc, k != 0);
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    const int q = warp & 3, hc = warp >> 2;
    const int i = i0 + q * 32 + lane;
    const bool row_ok = i < row_end;
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f};
    int Li[PtPad];
    float* myTrow = s_Trow + (size_t)(warp * 32 + lane) * TS;
    float* myTcol = s_Tcol + (size_t)(warp * 32 + lane) * TS;
    uint32_t* myX = s_X + (size_t)warp * kXWords;
#pragma unroll
    for (int k = 0; k < PtPad; ++k) {
      Li[k] = row_ok ? (__ldg(p.lfix + (size_t)i * PtPad + k) & ~31) : (kZeroFix * 32);
      myTrow[k] = 0.f;
      myTcol[k] = 0.f;
    }
    const float di = row_ok ? __ldg(p.d + i) : 0.f;
    const double e = p.e[b];
    const float k1 = (float)(1.0 / e), k2 = (float)(1.0 / (e * e));
    const float nx2scale = -2.f * 1.4426950408889634f * p.inv_temp;
    const float gscale4 = -800.f * p.inv_temp;                          // 4 * 2 * (-100) / temp
    const int cx = lane & 15, hx = lane >> 4;                           // column-phase role of this lane
    for (int it = 0; it < ntiles; ++it) {
      const int stage = it % kStages, acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kPN + hc * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      const int jbase = row_begin + (t0 + it) * kPN + hc * 16;
      mbar_wait(&full[stage], (it / kStages) & 1);
      const uint8_t* st = s_stage + (size_t)stage * kStageBytes + kPBBytes;
      const int4* sL = reinterpret_cast<const int4*>(st) + (size_t)hc * 16 * (PtPad / 4);
      const float dj_lane = (jbase + cx < row_end) ? __ldg(p.d + jbase + cx) : 0.f;

      auto chains = [&](int jj, int (&mg0)[2], int (&mg1)[2]) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int4* lj = sL + (jj + c) * (PtPad / 4);
          int ch[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
          for (int qd = 0; qd < NQ1; ++qd) {
            const int4 w = lj[qd];
            ch[0] = __viaddmax_s32(Li[4 * qd + 0], w.x, ch[0]);
            ch[1] = __viaddmax_s32(Li[4 * qd + 1], w.y, ch[1]);
            ch[2] = __viaddmax_s32(Li[4 * qd + 2], w.z, ch[2]);
            ch[3] = __viaddmax_s32(Li[4 * qd + 3], w.w, ch[3]);
          }
          mg0[c] = max(__vimax3_s32(ch[0], ch[1], ch[2]), ch[3]);
          int cg[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
#pragma unroll
          for (int qd = 0; qd < NQ2; ++qd) {
            const int4 w = lj[NQ1 + qd];
            cg[0] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 0], w.x, cg[0]);
            cg[1] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 1], w.y, cg[1]);
            cg[2] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 2], w.z, cg[2]);
            cg[3] = __viaddmax_s32(Li[4 * (NQ1 + qd) + 3], w.w, cg[3]);
          }
          mg1[c] = max(__vimax3_s32(cg[0], cg[1], cg[2]), cg[3]);
        }
      };
      auto tail = [&](int jj, const int (&mg0)[2], const int (&mg1)[2]) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int j = jbase + jj + c;
          const bool ok = row_ok && j < row_end && j >= i;                 // upper triangle + diagonal only
          const float sym = j > i ? 2.f : 1.f;                             // an unordered pair stands for (i,j) and (j,i)
          const float a = (ok && j != i) ? fmaxf(__uint_as_float(v[jj + c]), 0.f) : 0.f;
          const float djv = __shfl_sync(0xffffffffu, dj_lane, jj + c);     // warp-uniform: never under a lane predicate
          const float dd = ok ? di * djv : 0.f;
          const float gw4 = gscale4 * (a * k1 - dd * k2);                  // 0 for masked pairs
          const float as = a * sym, ds = dd * sym;
#pragma unroll
          for (int grp = 0; grp < (NQ2 ? 2 : 1); ++grp) {
            const int m = grp ? mg1[c] : mg0[c];
            const int pl = m & 31;
            const float u = ex2_approx((float)m * (1.f / (float)(1 << (kLogShift + 5))));
            const float e2 = ex2_approx(u * nx2scale);
            const float r = rcp_approx(1.f + e2);
            const float t1 = e2 * r;
            const float delta = r - t1;
            s1[grp] = fmaf(as, delta, s1[grp]);
            s2[grp] = fmaf(ds, delta, s2[grp]);
            const float t = (gw4 * t1) * (r * u);
            myTrow[pl + (grp ? 4 * NQ1 : 0)] += t;
            // column side: (t | p*) to the transpose buffer; the diagonal pair has no mirror image
            const uint32_t word = j > i ? ((__float_as_uint(t) & ~31u) | (uint32_t)pl) : (uint32_t)pl;
            myX[(grp * 16 + jj + c) * 33 + lane] = word;
          }
        }
      };
      int ma0[2], ma1[2];
      chains(0, ma0, ma1);
#pragma unroll
      for (int jj = 0; jj < 16; jj += 2) {
        int mb0[2] = {INT_MIN, INT_MIN}, mb1[2] = {INT_MIN, INT_MIN};
        if (jj + 2 < 16) chains(jj + 2, mb0, mb1);
        tail(jj, ma0, ma1);
        ma0[0] = mb0[0]; ma0[1] = mb0[1]; ma1[0] = mb1[0]; ma1[1] = mb1[1];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      // ---- column phase: lane (cx, hx) folds rows hx*16..+15 of column cx into its private accumulators ----
#pragma unroll
      for (int grp = 0; grp < (NQ2 ? 2 : 1); ++grp) {
        const uint32_t* col = myX + (grp * 16 + cx) * 33 + hx * 16;
#pragma unroll 4
        for (int r = 0; r < 16; ++r) {
          const uint32_t w = col[r];
          myTcol[(w & 31u) + (grp ? 4 * NQ1 : 0)] += __uint_as_float(w & ~31u);
        }
      }
      // ---- combine the 8 partial accumulators of every column inside the CTA, one RED per non-zero ----
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      {
        const int jt0 = row_begin + (t0 + it) * kPN;
        for (int ent = threadIdx.x; ent < kPN * PtPad; ent += kEpiWarps * 32) {
          const int c32 = ent / PtPad, k = ent - c32 * PtPad;
          float* base = s_Tcol + (size_t)(((c32 >> 4) * 4) * 32 + (c32 & 15)) * TS + k;   // warp (hc*4+q), lane hx*16+cx
          float tsum = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float* ptr = base + (size_t)(qq * 32 + hh * 16) * TS;
              tsum += *ptr;
              *ptr = 0.f;
            }
          if (tsum != 0.f && jt0 + c32 < row_end) atomicAdd(p.T + (size_t)(jt0 + c32) * PtPad + k, tsum);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      __syncwarp();                                   // X is rewritten by the next tile
    }
    if (row_ok) {
#pragma unroll 4
      for (int k = 0; k < PtPad; ++k) {
        const float t = myTrow[k];
        if (t != 0.f) atomicAdd(p.T + (size_t)i * PtPad + k, t);
      }
    }
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      const float a1 = warp_sum(s1[grp]), a2 = warp_sum(s2[grp]);
      if (lane == 0) { s_red[warp * 4 + grp * 2] = a1; s_red[warp * 4 + grp * 2 + 1] = a2; }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
    if (threadIdx.x < 4) {
      double t = 0.0;
      for (int w = 0; w < kEpiWarps; ++w) t += (double)s_red[w * 4 + threadIdx.x];
      atomicAdd(p.s + (size_t)b * 4 + threadIdx.x, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ------------------------------------------------------------------------------------------
// finish: dC = T / C (C = exp2(L)), dchat_p = sum_i dC_ip xh_i ; loss per group
//   grid (row chunks, B), 256 threads = features; dchat accumulated with atomicAdd
// ------------------------------------------------------------------------------------------
struct FinishParams {
  const bf16* h;
  const float* invn;
  const int* lfix;
  const float* T;
  const int* cu;
  const double* s;
  const double* e;
  float* dchat;       // (B, Pt, 256), zeroed by the launcher
  float* loss;        // (B, 2)
  int P1, P2, P1pad, PtPad, rows_per_cta;
};

template <int PTPAD>
__global__ void __launch_bounds__(256) modularity_finish_kernel(const FinishParams p) {
  __shared__ float s_dc[32][PTPAD];
  const int b = blockIdx.y, f = threadIdx.x;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int Pt = p.P1 + p.P2;
  if (blockIdx.x == 0 && f < 2) {
    const double e = p.e[b];
    const double s1 = p.s[(size_t)b * 4 + f * 2], s2 = p.s[(size_t)b * 4 + f * 2 + 1];
    p.loss[b * 2 + f] = (float)(-100.0 * (s1 / e - s2 / (e * e)));          // utils.py:222-228
  }
  const int r0 = row_begin + blockIdx.x * p.rows_per_cta;
  const int r1 = min(row_end, r0 + p.rows_per_cta);
  if (r0 >= r1) return;
  float acc[PTPAD];
#pragma unroll
  for (int k = 0; k < PTPAD; ++k) acc[k] = 0.f;
  for (int rb = r0; rb < r1; rb += 32) {
    const int nr = min(32, r1 - rb);
    __syncthreads();
    for (int idx = f; idx < 32 * PTPAD; idx += 256) {
      const int r = idx / PTPAD, k = idx % PTPAD;
      float v = 0.f;
      if (r < nr) {
        const size_t o = (size_t)(rb + r) * PTPAD + k;
        const float t = p.T[o];
        if (t != 0.f) {
          const float c = ex2_approx((float)(p.lfix[o] >> 5) * (1.f / (float)(1 << kLogShift)));
          v = c > 0.f ? t / c : 0.f;                   // relu gate: C == 0 never wins the max with u > 0
        }
      }
      s_dc[r][k] = v;
    }
    __syncthreads();
    for (int r = 0; r < nr; ++r) {
      const float xv = __bfloat162float(p.h[(size_t)(rb + r) * kD + f]) * __ldg(p.invn + rb + r);
#pragma unroll
      for (int k = 0; k < PTPAD; ++k) acc[k] += s_dc[r][k] * xv;
    }
  }
#pragma unroll
  for (int k = 0; k < PTPAD; ++k) {
    int tok = -1;
    if (k < p.P1) tok = k;
    else if (k >= p.P1pad && k - p.P1pad < p.P2) tok = p.P1 + k - p.P1pad;
    if (tok >= 0 && acc[k] != 0.f) atomicAdd(p.dchat + ((size_t)b * Pt + tok) * kD + f, acc[k]);
  }
}

int pad4(int v) { return (v + 3) & ~3; }
int quads1(int P1) { return P1 <= 8 ? 2 : (P1 <= 16 ? 4 : 8); }
int quads2(int P2) { return P2 == 0 ? 0 : (P2 <= 8 ? 2 : (P2 <= 16 ? 4 : 8)); }

template <int MODE, int NQ1, int NQ2>
int run_gram(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = gram_smem<MODE, NQ1, NQ2>();
  static_assert(smem <= 227 * 1024, "modularity_gram shared memory");
  static bool done = false;
  if (!done) {
    IMP_CUDA(cudaFuncSetAttribute(modularity_gram_kernel<MODE, NQ1, NQ2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  IMP_LAUNCH(MODE ? "modularity_gram_main" : "modularity_gram_degrees", st, modularity_gram_kernel<MODE, NQ1, NQ2><<<grid, kThreads, smem, st>>>(ta, tb, p));
  return IMP_OK;
}

template <int NQ1, int NQ2>
int run_pairs(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = pairs_smem<NQ1, NQ2>();
  static_assert(smem <= 227 * 1024, "modularity_pairs shared memory");
  static bool done = false;
  if (!done) {
    IMP_CUDA(cudaFuncSetAttribute(modularity_pairs_kernel<NQ1, NQ2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  IMP_LAUNCH("modularity_pairs", st, modularity_pairs_kernel<NQ1, NQ2><<<grid, kThreads, smem, st>>>(ta, tb, p));
  return IMP_OK;
}

struct Carve {
  bf16* xh; float* invn; int* lfix; float* d; float* T; double* e; double* s;
  size_t zero_off, zero_bytes, total;
};
Carve carve(void* ws, int total_rows, int B, int PtPad) {
  const size_t rpad = (size_t)total_rows + kBM;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  Carve c;
  c.xh = reinterpret_cast<bf16*>(base + off); off += up(rpad * kD * 2);
  c.invn = reinterpret_cast<float*>(base + off); off += up(rpad * 4);
  c.lfix = reinterpret_cast<int*>(base + off); off += up(rpad * PtPad * 4);
  c.zero_off = off;
  c.d = reinterpret_cast<float*>(base + off); off += up(rpad * 4);
  c.T = reinterpret_cast<float*>(base + off); off += up(rpad * PtPad * 4);
  c.e = reinterpret_cast<double*>(base + off); off += up((size_t)B * 8);
  c.s = reinterpret_cast<double*>(base + off); off += up((size_t)B * 4 * 8);
  c.zero_bytes = off - c.zero_off;
  c.total = off;
  return c;
}

}  // namespace

size_t modularity_workspace_bytes(int total_rows, int B, int P1, int P2) {
  const int PtPad = 4 * (quads1(P1) + quads2(P2));
  return carve(nullptr, total_rows, B, PtPad).total + 256;
}

int launch_modularity(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* chat, int P1, int P2,
                      float temp, void* workspace, float* loss, float* dchat, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  if (P1 < 1 || P1 > 32 || P2 < 0 || P2 > 32) IMP_FAIL(IMP_ERR_ARG, "modularity: token groups (%d,%d) must be in [1,32] and [0,32]", P1, P2);
  if (total_rows <= 0 || max_len <= 0) IMP_FAIL(IMP_ERR_ARG, "modularity: empty input");
  if (!(temp > 0.f)) IMP_FAIL(IMP_ERR_ARG, "modularity: temp must be positive");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) IMP_FAIL(IMP_ERR_ARG, "modularity: workspace must be 256-byte aligned");
  const int nq1 = quads1(P1), nq2 = quads2(P2);
  const int P1pad = 4 * nq1, PtPad = 4 * (nq1 + nq2), Pt = P1 + P2;
  Carve c = carve(workspace, total_rows, B, PtPad);
  // padding rows of xh / lfix are read by the tile loads of the last row block: keep them finite
  IMP_CUDA(cudaMemsetAsync(c.xh + (size_t)total_rows * kD, 0, (size_t)kBM * kD * 2, st));
  IMP_CUDA(cudaMemsetAsync(c.lfix + (size_t)total_rows * PtPad, 0, (size_t)kBM * PtPad * 4, st));
  IMP_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(workspace) + c.zero_off, 0, c.zero_bytes, st));
  IMP_CUDA(cudaMemsetAsync(dchat, 0, (size_t)B * Pt * kD * 4, st));

  PrepParams pp;
  pp.h = h; pp.cu = cu; pp.chat = chat; pp.xh = c.xh; pp.invn = c.invn; pp.lfix = c.lfix;
  pp.B = B; pp.P1 = P1; pp.P2 = P2; pp.P1pad = P1pad; pp.PtPad = PtPad;
  const size_t prep_smem = (size_t)PtPad * kD * 4;
  static bool prep_attr = false;
  if (!prep_attr) {
    IMP_CUDA(cudaFuncSetAttribute(modularity_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * kD * 4));
    prep_attr = true;
  }
  const int prep_chunks = std::max(1, std::min((max_len + 7) / 8, (4 * imp_num_sms() + B - 1) / B));
  IMP_LAUNCH("modularity_prep", st, modularity_prep_kernel<<<dim3(prep_chunks, B), 256, prep_smem, st>>>(pp));

  CUtensorMap ta, tb;
  int rc;
  const uint64_t rpad = (uint64_t)total_rows + kBM;
  if ((rc = imp_make_tmap_2d(&ta, c.xh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rpad, kD * 2, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = imp_make_tmap_2d(&tb, c.xh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rpad, kD * 2, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  GramParams gp;
  gp.cu = cu; gp.lfix = c.lfix; gp.d = c.d; gp.e = c.e; gp.T = c.T; gp.s = c.s; gp.inv_temp = 1.f / temp;
  const int row_blocks = (max_len + kBM - 1) / kBM;
  const int col_tiles = (max_len + kBN - 1) / kBN;
  // enough CTAs for >= 2 waves; at least 16 column tiles per CTA to amortise the A block load
  int nsplit = std::max(1, std::min((2 * imp_num_sms() + row_blocks * B - 1) / (row_blocks * B), std::max(1, col_tiles / 16)));
  gp.tiles_per_split = (col_tiles + nsplit - 1) / nsplit;
  nsplit = (col_tiles + gp.tiles_per_split - 1) / gp.tiles_per_split;
  const dim3 grid(row_blocks, nsplit, B);
  if ((rc = run_gram<0, 2, 0>(ta, tb, gp, grid, st))) return rc;
  CUtensorMap tp;
  if ((rc = imp_make_tmap_2d(&tp, c.xh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rpad, kD * 2, 64, kPN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  // pair sweep: triangular work per row block; split the column range so that >= 2 waves of CTAs exist
  int psplit = std::max(1, std::min((2 * imp_num_sms() + row_blocks * B - 1) / (row_blocks * B), std::max(1, col_tiles / 8)));
  const dim3 pgrid(B, psplit, row_blocks);
  static const bool symmetric = []() { const char* e = getenv("IMP_MODULARITY_SYMMETRIC"); return e ? atoi(e) != 0 : false; }();   // measured on B200: the full sweep is (slightly) faster, see DESIGN.md 5
#define IMP_GRAM(a, b2) rc = symmetric ? run_pairs<a, b2>(ta, tp, gp, pgrid, st) : run_gram<1, a, b2>(ta, tb, gp, grid, st)
  if (nq2 == 0) { if (nq1 == 2) IMP_GRAM(2, 0); else if (nq1 == 4) IMP_GRAM(4, 0); else IMP_GRAM(8, 0); }
  else if (nq2 == 2) { if (nq1 == 2) IMP_GRAM(2, 2); else if (nq1 == 4) IMP_GRAM(4, 2); else IMP_GRAM(8, 2); }
  else IMP_FAIL(IMP_ERR_ARG, "modularity: second token group supports at most 8 tokens (got %d)", P2);
#undef IMP_GRAM
  if (rc) return rc;

  FinishParams fp;
  fp.h = h; fp.invn = c.invn; fp.lfix = c.lfix; fp.T = c.T; fp.cu = cu; fp.s = c.s; fp.e = c.e;
  fp.dchat = dchat; fp.loss = loss; fp.P1 = P1; fp.P2 = P2; fp.P1pad = P1pad; fp.PtPad = PtPad;
  const int fin_chunks = std::max(1, std::min((max_len + 255) / 256, (4 * imp_num_sms() + B - 1) / B));
  fp.rows_per_cta = ((max_len + fin_chunks - 1) / fin_chunks + 31) & ~31;
  const dim3 fgrid((max_len + fp.rows_per_cta - 1) / fp.rows_per_cta, B);
  switch (PtPad) {
    case 8: IMP_LAUNCH("modularity_finish", st, modularity_finish_kernel<8><<<fgrid, 256, 0, st>>>(fp)); break;
    case 16: IMP_LAUNCH("modularity_finish", st, modularity_finish_kernel<16><<<fgrid, 256, 0, st>>>(fp)); break;
    case 24: IMP_LAUNCH("modularity_finish", st, modularity_finish_kernel<24><<<fgrid, 256, 0, st>>>(fp)); break;
    case 32: IMP_LAUNCH("modularity_finish", st, modularity_finish_kernel<32><<<fgrid, 256, 0, st>>>(fp)); break;
    case 40: IMP_LAUNCH("modularity_finish", st, modularity_finish_kernel<40><<<fgrid, 256, 0, st>>>(fp)); break;
    default: IMP_FAIL(IMP_ERR_ARG, "modularity: unsupported padded token count %d", PtPad);
  }
  IMP_LAUNCH_CHECK();
  return IMP_OK;
}
