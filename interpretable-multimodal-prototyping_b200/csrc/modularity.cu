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
#include <type_traits>
#include <limits.h>

namespace {

constexpr int kD = 256;
constexpr int kBM = 128;                 // rows (i) per CTA
constexpr int kBN = 64;                  // columns (j) per tile
constexpr int kABytes = kBM * kD * 2;    // 64 KB, four [128][64] swizzled boxes
constexpr int kBBytes = kBN * kD * 2;    // 32 KB, four [64][64] boxes
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kStages = 2;
// Fixed-point log-assignments: N = round((kCOff - log2 C) * 2^kLogShift) in [0, kNMax]  (C <= 2^kCOff = 16 by
// Cauchy-Schwarz: |xh| = 1, |chat_p| <= sqrt(256)); C == 0 (or below 2^-28) is kNMax.  Both operands are stored
// as integer-valued FLOATS: the column operand is (N << 5) | token index, the row operand 2^23 + (N << 5).  Their
// fp32 sum is exact (an integer below 2^24) and its bit pattern is kMagic + ((N_i + N_j) << 5 | p): the minimum
// over tokens yields the winning log-product and the arg-min in one word, with no int<->float conversion.
constexpr int kLogShift = 12;
constexpr int kNMax = (1 << 17) - 1;
constexpr int kCOff = 4;
constexpr int kMagic = 0x4B000000;
constexpr float kArgOff = 64.f + 2.f * kCOff;   // undoes 2^23 * 2^-(kLogShift+5) and the two offsets

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
  float* lfix;              // (Rpad, PtPad) integer-valued floats
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
        int fix = kNMax;
        if (real && dot > 0.f) fix = min(kNMax, max(0, __float2int_rn(((float)kCOff - log2f(dot)) * (float)(1 << kLogShift))));
        // low 5 bits: token index inside its group (column-side operand); the row side masks them off
        p.lfix[(size_t)row * p.PtPad + slot] = (float)(fix * 32 + (slot < p.P1pad ? slot : slot - p.P1pad));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Degrees of a general (signed) patch graph:  d_i = sum_{j != i} relu(xh_i . xh_j),  e = sum_i d_i.
//   CTA = 128 rows x a range of 64-column tiles; A block resident in smem, B tiles by TMA (2 stages),
//   tcgen05 128x64x256 into two TMEM buffers; 8 epilogue warps, thread = (row, 32 of the 64 columns).
// ------------------------------------------------------------------------------------------
struct GramParams {
  const int* cu;
  const float* lfix;        // (Rpad, PtPad)
  float* d;                 // (Rpad) degrees
  double* e;                // (B)
  float* T;                 // (Rpad, PtPad) atomicAdd
  double* s;                // (B, 2 groups): sum_ij (A_ij/e - d_i d_j/e^2) delta_ij
  const int* nonneg;        // (1) device flag: 1 when every element of h is >= 0 (closed-form degrees were used)
  int tiles_per_split;      // column tiles per CTA
  float inv_temp;
};

constexpr size_t kDegSmem = 1024 + kABytes + kStages * kBBytes + 256;

__global__ void __launch_bounds__(kThreads, 1)
modularity_degrees_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                          const GramParams p) {
  if (p.nonneg && __ldg(p.nonneg)) return;            // degrees came from the closed form (see prep)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_a = smem;
  uint8_t* s_stage = s_a + kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_stage + kStages * kBBytes);
  uint64_t* full = bars;                  // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;       // [kStages]  MMA commit -> TMA
  uint64_t* tfull = bars + 2 * kStages;   // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;           // [2] epilogue -> MMA
  uint64_t* afull = tempty + 2;           // A block landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);   // [8] block reduction scratch

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
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
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
    if (lane == 0) {
      mbar_arrive_expect_tx(afull, kABytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_a + bx * (kBM * 128), &tm_a, afull, bx * 64, i0);
      for (int it = 0; it < ntiles; ++it) {
        const int stage = it % kStages;
        mbar_wait_idle(&empty[stage], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[stage], kBBytes);
        uint8_t* dst = s_stage + (size_t)stage * kBBytes;
        const int j0 = row_begin + (t0 + it) * kBN;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tma_load_2d(dst + bx * (kBN * 128), &tm_b, &full[stage], bx * 64, j0);
      }
    }
  } else if (warp == kEpiWarps + 1) {
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
        const uint32_t sb = smem_u32(s_stage + (size_t)stage * kBBytes);
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
    const int q = warp & 3, hc = warp >> 2;
    const int i = i0 + q * 32 + lane;
    const bool row_ok = i < row_end;
    float acc_d = 0.f;
    for (int it = 0; it < ntiles; ++it) {
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + hc * 32, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);      // accumulator is in registers: release TMEM early
      const int jt0 = row_begin + (t0 + it) * kBN;
      const int jbase = jt0 + hc * 32;
      // interior tiles (no diagonal, no bag end) need no per-pair masks: relu + add only
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
    }
    if (row_ok) atomicAdd(p.d + i, acc_d);
    float tot = warp_sum(row_ok ? acc_d : 0.f);
    if (lane == 0) s_red[warp] = tot;
    asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < kEpiWarps; ++w) t += (double)s_red[w];
      atomicAdd(p.e + b, t);
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
// Pair sweep (traces + T): the dominant kernel of the training step.
//   CTA = 128 rows x a range of 64-column tiles; 16 warps, thread = (row i, 16 of the 64 columns), 128 registers
//   per thread (a 17th warp would cap every thread at 96).  There are no dedicated producer / MMA warps: the
//   pair work of a tile is ~3 us, so lane 0 of warp 0 issues the asynchronous work inline -- at the top of
//   tile t the TMA loads of B(t+1) and of the L/degree tile t+2, halfway through tile t the tcgen05 MMA of
//   tile t+1 (128x64x256 into the other TMEM buffer).  Per pair every thread runs
//     - the (max,x) contraction over tokens as one VIADDMNMX (min form) per (pair, token) on the ALU pipe,
//       four independent chains per token group;  the row operand carries the float "magic" offset
//       0x4B000000, so the winning sum IS the bit pattern of the float 2^23 + n and needs no I2F;
//     - the tanh / gradient tail on packed fp32x2 instructions (FFMA2/FMUL2) over two columns at a time;
//     - T_i[p*] += t in a thread-private, token-major shared array (bank = thread: conflict free).
// ------------------------------------------------------------------------------------------
constexpr int kSwWarps = 16;
constexpr int kSwEpi = kSwWarps * 32;               // 512 epilogue threads
constexpr int kSwThreads = kSwEpi;
constexpr int kLStages = 4;
constexpr int kDTile = (kBN + 4) * 4;               // degree tile: 64 floats from a 16-byte aligned start

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)), "l"(*reinterpret_cast<uint64_t*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
  return *reinterpret_cast<float2*>(&d);
}

template <int NQ1, int NQ2>
constexpr size_t sweep_smem() {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  return 1024 + kABytes + kBBytes + (size_t)kLStages * (kBN * PtPad * 4 + kDTile) + (size_t)PtPad * kSwEpi * 4 + 512;
}

template <int NQ1, int NQ2>
__global__ void __launch_bounds__(kSwThreads, 1)
modularity_sweep_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                        const GramParams p) {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  constexpr int NG = NQ2 ? 2 : 1;
  constexpr int kLBytes = kBN * PtPad * 4;
  constexpr int kLStage = kLBytes + kDTile;             // L tile + degree tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_a = smem;
  uint8_t* s_b = s_a + kABytes;
  uint8_t* s_l = s_b + kBBytes;
  float* s_T = reinterpret_cast<float*>(s_l + kLStages * kLStage);       // [PtPad][512]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_T + (size_t)PtPad * kSwEpi);
  uint64_t* bfull = bars;                 // TMA -> MMA
  uint64_t* bempty = bars + 1;            // MMA commit -> TMA
  uint64_t* tfull = bars + 2;             // [2] MMA -> epilogue
  uint64_t* tempty = bars + 4;            // [2] epilogue -> MMA
  uint64_t* lfull = bars + 6;             // [kLStages] TMA -> epilogue
  uint64_t* lempty = lfull + kLStages;    // [kLStages] 16 epilogue warps -> TMA
  uint64_t* afull = lempty + kLStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);                 // [16][4]

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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    mbar_init(bfull, 1); mbar_init(bempty, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kSwWarps); }
    for (int i = 0; i < kLStages; ++i) { mbar_init(&lfull[i], 1); mbar_init(&lempty[i], kSwWarps); }
    mbar_init(afull, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool driver = (warp == 0 && lane == 0);       // issues TMA and MMA for the whole CTA

  auto load_b = [&](int it) {                          // B box of tile `it` (single stage: after the MMA of tile it-1)
    mbar_wait(bempty, (it & 1) ^ 1);
    mbar_arrive_expect_tx(bfull, kBBytes);
    const int j0 = row_begin + (t0 + it) * kBN;
#pragma unroll
    for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_b + bx * (kBN * 128), &tm_b, bfull, bx * 64, j0);
  };
  auto load_l = [&](int it) {                          // L tile + degree tile of tile `it`
    const int ls = it % kLStages;
    mbar_wait(&lempty[ls], ((it / kLStages) & 1) ^ 1);
    mbar_arrive_expect_tx(&lfull[ls], kLStage);
    const int j0 = row_begin + (t0 + it) * kBN;
    uint8_t* dst = s_l + (size_t)ls * kLStage;
    bulk_load(dst, p.lfix + (size_t)j0 * PtPad, kLBytes, &lfull[ls]);
    bulk_load(dst + kLBytes, p.d + (j0 & ~3), kDTile, &lfull[ls]);        // bulk copies need 16-byte aligned sources
  };
  auto issue_mma = [&](int it) {
    constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
    const int acc = it & 1;
    mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
    mbar_wait(bfull, it & 1);
    tc_fence_after();
    const uint32_t sa = smem_u32(s_a), sb = smem_u32(s_b);
#pragma unroll
    for (int k = 0; k < kD / 16; ++k) {
      const uint64_t ad = umma_desc_sw128(sa + (k >> 2) * (kBM * 128) + (k & 3) * 32, 0, 1024);
      const uint64_t bd = umma_desc_sw128(sb + (k >> 2) * (kBN * 128) + (k & 3) * 32, 0, 1024);
      umma_f16(tmem_base + acc * kBN, ad, bd, idesc, k != 0);
    }
    umma_commit(bempty);
    umma_commit(&tfull[acc]);
  };
  if (driver) {
    mbar_arrive_expect_tx(afull, kABytes);
#pragma unroll
    for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_a + bx * (kBM * 128), &tm_a, afull, bx * 64, i0);
    load_b(0);
    load_l(0);
    if (ntiles > 1) load_l(1);
    mbar_wait(afull, 0);
    issue_mma(0);
  }
  {
    // ------------------------------ epilogue: thread = (row i, 16 of the 64 columns) ------------------
    const int q = warp & 3, hc = warp >> 2;
    const int tid = threadIdx.x;                       // = hc*128 + row in block
    const int i = i0 + q * 32 + lane;
    const bool row_ok = i < row_end;
    float* myT = s_T + tid;                            // element p at myT[p * 512]
    float2 Li[PtPad / 2];                              // row operand, token pairs: 2^23 + (N_i << 5)
#pragma unroll
    for (int k = 0; k < PtPad; k += 2) {
      const float2 w = row_ok ? __ldg(reinterpret_cast<const float2*>(p.lfix + (size_t)i * PtPad + k))
                              : make_float2((float)(kNMax << 5), (float)(kNMax << 5));
      Li[k >> 1] = make_float2((float)((int)w.x & ~31) + 8388608.f, (float)((int)w.y & ~31) + 8388608.f);
      myT[k * kSwEpi] = 0.f;
      myT[(k + 1) * kSwEpi] = 0.f;
    }
    const float di = row_ok ? __ldg(p.d + i) : 0.f;
    const double e = p.e[b];
    const float k1 = (float)(1.0 / e), k2 = (float)(1.0 / (e * e));
    const float gs4 = -800.f * p.inv_temp;                                 // 4 * 2 * (-100) / temp
    const float2 c0 = make_float2(gs4 * k1, gs4 * k1);
    const float2 nci = make_float2(-gs4 * k2 * di, -gs4 * k2 * di);
    const float nx2 = -2.f * 1.4426950408889634f * p.inv_temp;             // tanh(u/temp) via exp2(-2 u log2e / temp)
    const float2 nx2s = make_float2(nx2, nx2);
    const float2 kscale = make_float2(-1.f / (float)(1 << (kLogShift + 5)), -1.f / (float)(1 << (kLogShift + 5)));
    const float2 koff = make_float2(kArgOff, kArgOff);
    const float2 one2 = make_float2(1.f, 1.f), neg2 = make_float2(-1.f, -1.f);
    float2 sg[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};

    for (int it = 0; it < ntiles; ++it) {
      const int acc = it & 1, ls = it % kLStages;
      if (driver) {
        if (it + 1 < ntiles) load_b(it + 1);
        if (it + 2 < ntiles) load_l(it + 2);
      }
      __syncwarp();
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[8];                                     // columns 0-7 now, 8-15 halfway through the tile
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kBN + hc * 16;
      tmem_ld8(taddr, v);
      tmem_ld_wait();
      mbar_wait(&lfull[ls], (it / kLStages) & 1);
      const uint8_t* st = s_l + (size_t)ls * kLStage;
      const float4* sL = reinterpret_cast<const float4*>(st) + (size_t)hc * 16 * (PtPad / 4);
      const int jt0 = row_begin + (t0 + it) * kBN;
      const float* sd = reinterpret_cast<const float*>(st + kLBytes) + (jt0 & 3) + hc * 16;
      const int jbase = jt0 + hc * 16;
      const bool interior = (i0 + kBM <= row_end) && (jt0 + kBN <= row_end) && (jt0 >= i0 + kBM || jt0 + kBN <= i0);

      // (min,+) contraction over tokens: one packed FADD2 (FMA pipe) per token PAIR, one 3-input FMNMX3
      // (ALU pipe) folding both sums into one of four independent chains per token group
      auto chains = [&](int jj, int (&m)[2][2]) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float4* lj = sL + (jj + c) * (PtPad / 4);
          float ch[4] = {3e38f, 3e38f, 3e38f, 3e38f};
#pragma unroll
          for (int qd = 0; qd < NQ1; ++qd) {
            const float4 w = lj[qd];
            const float2 sa = add2(Li[2 * qd], make_float2(w.x, w.y));
            const float2 sb = add2(Li[2 * qd + 1], make_float2(w.z, w.w));
            ch[(2 * qd) & 3] = fminf(fminf(sa.x, sa.y), ch[(2 * qd) & 3]);
            ch[(2 * qd + 1) & 3] = fminf(fminf(sb.x, sb.y), ch[(2 * qd + 1) & 3]);
          }
          m[c][0] = __float_as_int(fminf(fminf(ch[0], ch[1]), fminf(ch[2], ch[3])));
          float cg[4] = {3e38f, 3e38f, 3e38f, 3e38f};
#pragma unroll
          for (int qd = 0; qd < NQ2; ++qd) {
            const float4 w = lj[NQ1 + qd];
            const float2 sa = add2(Li[2 * (NQ1 + qd)], make_float2(w.x, w.y));
            const float2 sb = add2(Li[2 * (NQ1 + qd) + 1], make_float2(w.z, w.w));
            cg[(2 * qd) & 3] = fminf(fminf(sa.x, sa.y), cg[(2 * qd) & 3]);
            cg[(2 * qd + 1) & 3] = fminf(fminf(sb.x, sb.y), cg[(2 * qd + 1) & 3]);
          }
          m[c][1] = NQ2 ? __float_as_int(fminf(fminf(cg[0], cg[1]), fminf(cg[2], cg[3]))) : 0;
        }
      };
      auto tail = [&](int jj, const int (&m)[2][2], auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        float2 a2 = make_float2(fmaxf(__uint_as_float(v[jj & 7]), 0.f), fmaxf(__uint_as_float(v[(jj & 7) + 1]), 0.f));
        float2 dj2 = make_float2(sd[jj], sd[jj + 1]);
        if (!INTERIOR) {
          const int j = jbase + jj;
          const bool ok0 = row_ok && j < row_end, ok1 = row_ok && j + 1 < row_end;
          a2.x = (ok0 && j != i) ? a2.x : 0.f;
          a2.y = (ok1 && j + 1 != i) ? a2.y : 0.f;
          dj2.x = ok0 ? dj2.x : 0.f;
          dj2.y = ok1 ? dj2.y : 0.f;
        }
        const float2 gw = fma2(a2, c0, mul2(dj2, nci));                 // 4*2*(-100)/temp * (A/e - d_i d_j/e^2); 0 when masked
#pragma unroll
        for (int grp = 0; grp < NG; ++grp) {
          const int m0 = m[0][grp], m1 = m[1][grp];
          const int pl0 = m0 & 31, pl1 = m1 & 31;
          const float2 F = make_float2(__int_as_float(m0 & ~31), __int_as_float(m1 & ~31));
          const float2 arg = fma2(F, kscale, koff);                     // log2(C_i C_j) of the winning token
          float2 u = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
          const float2 x = mul2(u, nx2s);
          const float2 e2 = make_float2(ex2_approx(x.x), ex2_approx(x.y));   // exp(-2u/temp)
          const float2 den = add2(e2, one2);
          const float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
          const float2 t1 = mul2(e2, r);
          const float2 delta = fma2(t1, neg2, r);                       // tanh(u/temp) = (1-e2)/(1+e2)
          sg[grp] = fma2(gw, delta, sg[grp]);                           // sum of gs4 * (A/e - d_i d_j/e^2) * delta
          const float2 t = mul2(mul2(gw, t1), mul2(r, u));              // 2 g (1-delta^2)/temp * u, 1-delta^2 = 4 e2 r^2
          float* T0 = myT + (size_t)(pl0 + (grp ? 4 * NQ1 : 0)) * kSwEpi;
          *T0 += t.x;
          float* T1 = myT + (size_t)(pl1 + (grp ? 4 * NQ1 : 0)) * kSwEpi;
          *T1 += t.y;
        }
      };
      auto sweep = [&](auto interior_tag) {
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
          int m[2][2];
          chains(jj, m);
          tail(jj, m, interior_tag);
          if (jj == 6) {                                 // second half of the accumulator row, then release the TMEM buffer
            tmem_ld8(taddr + 8, v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (driver && it + 1 < ntiles) issue_mma(it + 1);
            __syncwarp();
          }
        }
      };
      if (interior) sweep(std::true_type{}); else sweep(std::false_type{});
      __syncwarp();
      if (lane == 0) mbar_arrive(&lempty[ls]);
    }
    // ---------------- flush ----------------
    {
      const float inv_gs4 = 1.f / gs4;
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        const float a1 = warp_sum((sg[grp].x + sg[grp].y) * inv_gs4);   // = sum A delta / e - sum d_i d_j delta / e^2
        if (lane == 0) s_red[warp * 2 + grp] = a1;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kSwEpi) : "memory");
    if (tid < 2) {
      double t = 0.0;
      for (int w = 0; w < kSwWarps; ++w) t += (double)s_red[w * 2 + tid];
      atomicAdd(p.s + (size_t)b * 2 + tid, t);
    }
    // the four column-quarter threads of a row are combined before one RED per (row, token)
    for (int idx = tid; idx < kBM * PtPad; idx += kSwEpi) {
      const int r = idx & (kBM - 1), k = idx >> 7;
      const float* src = s_T + (size_t)k * kSwEpi + r;
      const float t = (src[0] + src[kBM]) + (src[2 * kBM] + src[3 * kBM]);
      if (t != 0.f && i0 + r < row_end) atomicAdd(p.T + (size_t)(i0 + r) * PtPad + k, t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------
// finish: dC = T / C (C = exp2(L)), dchat_p = sum_i dC_ip xh_i ; loss per group
//   grid (row chunks, B), 256 threads = features; dchat accumulated with atomicAdd
// ------------------------------------------------------------------------------------------
struct FinishParams {
  const bf16* h;
  const float* invn;
  const float* lfix;
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
    p.loss[b * 2 + f] = (float)(-100.0 * p.s[(size_t)b * 2 + f]);          // utils.py:222-228: -100 tr((W/e) delta)
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
          const int nfix = (int)p.lfix[o] >> 5;
          const float c = ex2_approx((float)kCOff - (float)nfix * (1.f / (float)(1 << kLogShift)));
          v = nfix < kNMax ? t / c : 0.f;              // relu gate: C == 0 never wins the max with u > 0
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

int run_degrees(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  static bool done = false;
  if (!done) {
    IMP_CUDA(cudaFuncSetAttribute(modularity_degrees_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDegSmem));
    done = true;
  }
  IMP_LAUNCH("modularity_gram_degrees", st, modularity_degrees_kernel<<<grid, kThreads, kDegSmem, st>>>(ta, tb, p));
  return IMP_OK;
}

template <int NQ1, int NQ2>
int run_sweep(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = sweep_smem<NQ1, NQ2>();
  static_assert(smem <= 227 * 1024, "modularity_sweep shared memory");
  static bool done = false;
  if (!done) {
    IMP_CUDA(cudaFuncSetAttribute(modularity_sweep_kernel<NQ1, NQ2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done = true;
  }
  IMP_LAUNCH("modularity_sweep", st, modularity_sweep_kernel<NQ1, NQ2><<<grid, kSwThreads, smem, st>>>(ta, tb, p));
  return IMP_OK;
}

struct Carve {
  bf16* xh; float* invn; float* lfix; float* d; float* T; double* e; double* s;
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
  c.lfix = reinterpret_cast<float*>(base + off); off += up(rpad * PtPad * 4);
  c.zero_off = off;
  c.d = reinterpret_cast<float*>(base + off); off += up(rpad * 4);
  c.T = reinterpret_cast<float*>(base + off); off += up(rpad * PtPad * 4);
  c.e = reinterpret_cast<double*>(base + off); off += up((size_t)B * 8);
  c.s = reinterpret_cast<double*>(base + off); off += up((size_t)B * 2 * 8);
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
  gp.nonneg = nullptr;
  if ((rc = run_degrees(ta, tb, gp, grid, st))) return rc;
#define IMP_SWEEP(a, b2) rc = run_sweep<a, b2>(ta, tb, gp, grid, st)
  if (nq2 == 0) { if (nq1 == 2) IMP_SWEEP(2, 0); else if (nq1 == 4) IMP_SWEEP(4, 0); else IMP_SWEEP(8, 0); }
  else if (nq2 == 2) { if (nq1 == 2) IMP_SWEEP(2, 2); else if (nq1 == 4) IMP_SWEEP(4, 2); else IMP_SWEEP(8, 2); }
  else IMP_FAIL(IMP_ERR_ARG, "modularity: second token group supports at most 8 tokens (got %d)", P2);
#undef IMP_SWEEP
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
