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
//   prep            : per 128-row tile: inv-norms, xh (bf16, the Gram operand), C = |x|^-1 h . chat^T on tcgen05 (h exact
//                     in bf16, chat as hi + lo bf16 parts) as fixed-point log2 in the tiled layout the sweep streams,
//                     column sums of xh, sign flag
//   degrees_closed  : all features >= 0 (always true behind path_net's ReLU): relu is the identity on the Gram
//                     matrix, so d_i = xh_i . (sum_j xh_j) - xh_i . xh_i  -- O(N D) instead of an N x N sweep
//   degrees (Gram)  : general signed features: tcgen05 Gram tiles 128 x 64 (K = 256), relu + row sums
//   sweep           : the trace and T_ip = sum_j 2 g_ij (1-delta^2)/temp u_ij [p = argmax]; Gram tiles on tcgen05,
//                     the (max,x) contraction over tokens in the log domain as a (min,+) contraction of
//                     integer-valued floats: FADD2 + FMNMX3 per token pair, arg-min in the low mantissa bits
//   finish          : dC = T / C, dchat_p = sum_i dC_ip xh_i as a tcgen05 GEMM (A = dC / |x| split hi + lo, B = h), loss per group
// Up to two token groups (<= 32 and <= 8 tokens) share A, d and e (the reference evaluates the
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
// fp32 sum is exact (an integer below 2^24) whose bit pattern is 0x4B000000 + ((N_i + N_j) << 5 | p): the minimum
// over tokens yields the winning log-product and the arg-min in one word, with no int<->float conversion.
// Layout of the column operand in HBM ("tiled"): [absolute 64-row tile][token slot][64 rows], so that one bulk
// copy brings a tile in the order the sweep reads it: for one token, 4 consecutive columns per 128-bit load.
constexpr int kLogShift = 12;
constexpr int kNMax = (1 << 17) - 1;
constexpr int kCOff = 4;
constexpr float kArgOff = 64.f + 2.f * kCOff;   // undoes 2^23 * 2^-(kLogShift+5) and the two offsets
__host__ __device__ __forceinline__ size_t lfix_index(int row, int slot, int PtPad) {
  return ((size_t)(row >> 6) * PtPad + slot) * 64 + (row & 63);
}

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
// non-blocking mbarrier probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) __nanosleep(40);
}
// T[slot] += v on the token-major fixed-point array (row stride 512 B): one IMAD for the address (the compiler's
// own form is shift + mask + add) and a fire-and-forget ATOMS.ADD
__device__ __forceinline__ void red_shared_add(uint32_t base, int slot, int v) {
  asm volatile(
      "{\n\t"
      ".reg .u32 a;\n\t"
      "mad.lo.u32 a, %1, 512, %0;\n\t"
      "red.shared.add.s32 [a], %2;\n\t"
      "}" ::"r"(base), "r"(slot), "r"(v)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
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

// ------------------------------------------------------------------------------------------
// prep: CTA = one absolute 64-row tile, 256 threads.
//   phase A  warp per row: norm, xh (bf16) and inv-norm to HBM, normalised fp32 row to smem, sign flag
//   phase B  per bag intersecting the tile: tokens to smem, C = xs . chat^T with a 2-row x TPG-token register
//            tile per thread (K = 256 in float4 steps), fixed-point log2 into the tiled layout; column sums of
//            the bf16-rounded xh per bag (closed-form degrees)
// ------------------------------------------------------------------------------------------
struct PrepParams {
  const bf16* h;            // (R,256)
  const int* cu;
  const float* chat;        // (B, Pt, 256): tokens normalised across tokens per feature
  bf16* xh;                 // (Rpad,256)
  float* invn;              // (Rpad)
  float* lfix;              // tiled, see lfix_index
  float* colsum;            // (B,256)  sum_j xh_j (bf16-rounded values)
  int* negflag;             // (1) set to 1 when any element of h is negative
  int R, B, P1, P2, P1pad;
  int row_lo, row_hi;       // this call covers global rows [row_lo, row_hi) (row_lo % 64 == 0); h points at row row_lo
};
// ------------------------------------------------------------------------------------------
// prep on tcgen05.  C = |x|^-1 (h . chat^T): h is EXACT in bf16 (it is stored that way), so with chat split into
// hi + lo bf16 parts (chat - hi - lo ~ 2^-17 chat) the tensor core delivers the inner products to ~4e-6 relative --
// far inside the 2^-12 log2 resolution of the fixed-point assignments -- as S = h [chat_hi ; chat_lo]^T (M 128,
// N = 2 PTPAD, K 256, fp32 accumulation in TMEM).  One CTA per 128-row tile (two CTAs resident per SM): TMA brings the
// h tile, the row threads take the norms from it while the MMAs run, the epilogue (thread = row = TMEM lane) writes the
// fixed-point log2 assignments in the tiled layout (coalesced: for one slot a warp writes 32 consecutive rows), then the
// tile is normalised in place and leaves as x_hat through coalesced stores; column sums of x_hat per bag by thread =
// feature.  The fp32-FMA version this replaces spent 0.66 ms per 32 bags (14 % of HBM), bound by shared-memory reads.
// ------------------------------------------------------------------------------------------
constexpr int kPtThreads = 5 * 32;               // 4 row warps + 1 issuing warp
template <int PTPAD>
constexpr size_t prep_tc_smem() { return 1024 + (size_t)kABytes + (size_t)2 * PTPAD * 512 + 128 * 4 + 64; }

template <int PTPAD>
__global__ void __launch_bounds__(kPtThreads, 2)
modularity_prep_tc_kernel(const __grid_constant__ CUtensorMap tm_h, const PrepParams p) {
  constexpr int NB2 = 2 * PTPAD;                   // rows of the B operand: chat_hi slots, then chat_lo slots
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_h = smem;                             // [128 rows][256] bf16, four swizzled boxes
  uint8_t* s_c = s_h + kABytes;                    // B operand: four boxes [NB2][64]
  float* s_inv = reinterpret_cast<float*>(s_c + (size_t)NB2 * 512);   // [128] 1 / |x_row|
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 128);
  uint64_t* full = bars;                           // TMA -> everyone
  uint64_t* sfull = bars + 1;                      // MMA -> epilogue (one phase per bag of the tile)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = p.row_lo + blockIdx.x * kBM;
  const int tile_end = min(r0 + kBM, p.row_hi);
  const int data_end = __ldg(p.cu + p.B);          // rows past cu[B] belong to no bag
  const int Pt = p.P1 + p.P2;

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_h);
      mbar_init(full, 1);
      mbar_init(sfull, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 4 && elect_one()) {
    mbar_arrive_expect_tx(full, kABytes);
#pragma unroll
    for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_h + bx * (kBM * 128), &tm_h, full, bx * 64, r0 - p.row_lo);
  }
  __syncwarp();

  // ---- norms of the rows (thread = row), straight from the swizzled tile ----
  const int n = threadIdx.x;                       // row inside the tile for warps 0-3
  const int row = r0 + n;
  float inv = 0.f;
  bool neg = false;
  if (warp < 4) {
    mbar_wait(full, 0);
    float ss = 0.f;
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
      const uint4 raw = *reinterpret_cast<const uint4*>(s_h + (c >> 3) * (kBM * 128) + n * 128 + (((c & 7) ^ (n & 7)) << 4));
      const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = bf16lo(wds[e]), hi = bf16hi(wds[e]);
        ss = fmaf(lo, lo, fmaf(hi, hi, ss));
        neg |= (lo < 0.f) || (hi < 0.f);
      }
    }
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);          // F.normalize eps (utils.py:179,193)
    s_inv[n] = inv;
    if (row < tile_end) p.invn[row] = inv;
    neg = neg && row < tile_end && row < data_end;
    if (__any_sync(0xffffffffu, neg) && lane == 0) *p.negflag = 1;
  }

  // ---- per bag of the tile: B = [chat_hi ; chat_lo], MMA, fixed-point assignments ----
  int b = find_segment(p.cu, p.B, r0);
  int phase = 0;
  for (; b < p.B; ++b) {
    const int bag_lo = __ldg(p.cu + b);
    if (bag_lo >= tile_end) break;
    const int sa = max(bag_lo, r0), sb = min(__ldg(p.cu + b + 1), tile_end);
    if (sa >= sb) continue;
    __syncthreads();                               // the previous bag's MMAs have been consumed (epilogue below ends with the TMEM reads)
    // four token rows in flight per thread: one float4 at a time left every CTA waiting 16 x an L2 round trip here
    // (a quarter of the kernel's stall samples)
    for (int i0 = threadIdx.x; i0 < PTPAD * 64; i0 += 4 * kPtThreads) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kPtThreads;
        const int slot = i >> 6, c = (i & 63) << 2;
        int src = -1;
        if (i < PTPAD * 64) {
          if (slot < p.P1) src = slot;
          else if (slot >= p.P1pad && slot - p.P1pad < p.P2) src = p.P1 + slot - p.P1pad;
        }
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) v[u] = __ldg(reinterpret_cast<const float4*>(p.chat + ((size_t)b * Pt + src) * kD + c));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kPtThreads;
        if (i >= PTPAD * 64) break;
        const int slot = i >> 6, c = (i & 63) << 2;
        const uint32_t h01 = pack_bf16x2(v[u].x, v[u].y), h23 = pack_bf16x2(v[u].z, v[u].w);
        const uint32_t l01 = pack_bf16x2(v[u].x - bf16lo(h01), v[u].y - bf16hi(h01)), l23 = pack_bf16x2(v[u].z - bf16lo(h23), v[u].w - bf16hi(h23));
        const uint32_t boxo = (uint32_t)((c >> 6) * (NB2 * 128));
        const uint32_t swz = (uint32_t)((c & 7) << 1);
        const int rh = slot, rl = PTPAD + slot;
        *reinterpret_cast<uint2*>(s_c + boxo + rh * 128 + ((((c & 63) >> 3) ^ (rh & 7)) << 4) + swz) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(s_c + boxo + rl * 128 + ((((c & 63) >> 3) ^ (rl & 7)) << 4) + swz) = make_uint2(l01, l23);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4) {
      mbar_wait(full, 0);
      tc_fence_after();
      if (elect_one()) {
        constexpr uint32_t idesc = umma_idesc_bf16(kBM, NB2, 0, 0);
        const uint32_t sh = smem_u32(s_h), sc = smem_u32(s_c);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t ad = umma_desc_sw128(sh + (k >> 2) * (kBM * 128) + (k & 3) * 32, 0, 1024);
          const uint64_t bd = umma_desc_sw128(sc + (k >> 2) * (NB2 * 128) + (k & 3) * 32, 0, 1024);
          umma_f16(tmem_base, ad, bd, idesc, k != 0);
        }
        umma_commit(sfull);
      }
      __syncwarp();
    } else {
      mbar_wait(sfull, phase);
      tc_fence_after();
      const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
      const bool mine = row >= sa && row < sb;
#pragma unroll 1
      for (int s0 = 0; s0 < PTPAD; s0 += 8) {
        uint32_t vh[8], vl[8];
        tmem_ld8(tmem_base + lane_addr + s0, vh);
        tmem_ld8(tmem_base + lane_addr + PTPAD + s0, vl);
        tmem_ld_wait();
        if (mine) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int slot = s0 + e;
            const float dot = (__uint_as_float(vh[e]) + __uint_as_float(vl[e])) * inv;
            int fix = kNMax;                        // padding slots have zero tokens: dot == 0
            if (dot > 0.f) fix = min(kNMax, max(0, __float2int_rn(((float)kCOff - log2f(dot)) * (float)(1 << kLogShift))));
            // low 5 bits: token index inside its group (column-side operand); the row side masks them off
            p.lfix[lfix_index(row, slot, PTPAD)] = (float)(fix * 32 + (slot < p.P1pad ? slot : slot - p.P1pad));
          }
        }
      }
      tc_fence_before();
    }
    phase ^= 1;
  }
  // rows of the tile that belong to no bag (past cu[B]: tile padding, or the unused tail of a buffer sized for the
  // worst case) are masked in the sweep, but its loads must stay finite: 0 * NaN would poison the sums
  if (warp < 4 && row >= data_end && row < ((p.R + kBM + kBN + 63) / 64) * 64) {
    for (int slot = 0; slot < PTPAD; ++slot) p.lfix[lfix_index(row, slot, PTPAD)] = (float)(kNMax * 32);
  }
  __syncthreads();                                 // every MMA that reads the h tile has retired (the epilogues waited on sfull)
  // ---- x_hat in place (bf16), column sums per bag, coalesced store ----
  if (warp < 4) {
    if (warp < 4) mbar_wait(full, 0);
#pragma unroll 4
    for (int c = 0; c < 32; ++c) {
      uint4* ptr = reinterpret_cast<uint4*>(s_h + (c >> 3) * (kBM * 128) + n * 128 + (((c & 7) ^ (n & 7)) << 4));
      const uint4 raw = *ptr;
      const uint32_t wds[4] = {raw.x, raw.y, raw.z, raw.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = pack_bf16x2(bf16lo(wds[e]) * inv, bf16hi(wds[e]) * inv);
      *ptr = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  if (warp < 4) {
    const int nvalid = max(0, tile_end - r0);
    for (int it = threadIdx.x; it < nvalid * 32; it += 128) {
      const int r = it >> 5, c = (it & 31) << 3;
      *reinterpret_cast<uint4*>(p.xh + (size_t)(r0 + r) * kD + c) =
          *reinterpret_cast<const uint4*>(s_h + (c >> 6) * (kBM * 128) + r * 128 + ((((c & 63) >> 3) ^ (r & 7)) << 4));
    }
    // column sums of the bf16-rounded unit rows, per bag (thread = 2 features)
    int bb = find_segment(p.cu, p.B, r0);
    for (; bb < p.B; ++bb) {
      const int bag_lo = __ldg(p.cu + bb);
      if (bag_lo >= tile_end) break;
      const int sa = max(bag_lo, r0), sb = min(__ldg(p.cu + bb + 1), tile_end);
      if (sa >= sb) continue;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int f = half * 128 + threadIdx.x;
        float acc = 0.f;
        for (int r = sa - r0; r < sb - r0; ++r)
          acc += __bfloat162float(*reinterpret_cast<const bf16*>(s_h + (f >> 6) * (kBM * 128) + r * 128 + ((((f & 63) >> 3) ^ (r & 7)) << 4) + ((f & 7) << 1)));
        atomicAdd(p.colsum + (size_t)bb * kD + f, acc);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------
// closed-form degrees for non-negative features: d_i = xh_i . S_b - xh_i . xh_i,  e_b = sum_i d_i
//   (relu is the identity on the Gram matrix when every xh >= 0; the diagonal is excluded, utils.py:194-196)
// ------------------------------------------------------------------------------------------
struct DegParams {
  const bf16* xh;
  const int* cu;
  const float* colsum;
  const int* negflag;
  float* d;
  double* e;
  int R, B;
};
__global__ void __launch_bounds__(256) modularity_degrees_closed_kernel(const DegParams p) {
  if (__ldg(p.negflag)) return;                     // signed features: the Gram sweep computes the degrees
  __shared__ float s_sum[8];
  __shared__ int s_bag[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 64 + warp * 8;
  const int data_end = __ldg(p.cu + p.B);
  int cur = -1;
  float acc = 0.f;
  float s[8];
  for (int rr = 0; rr < 8; ++rr) {
    const int row = r0 + rr;
    if (row >= p.R || row >= data_end) break;          // rows past cu[B] belong to no bag
    int b = cur;
    if (cur < 0 || row >= __ldg(p.cu + cur + 1)) b = find_segment(p.cu, p.B, row);
    if (b != cur) {
      if (cur >= 0 && lane == 0) atomicAdd(p.e + cur, (double)acc);
      acc = 0.f;
      cur = b;
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(p.colsum + (size_t)b * kD + lane * 8));
      const float4 s1 = __ldg(reinterpret_cast<const float4*>(p.colsum + (size_t)b * kD + lane * 8 + 4));
      s[0] = s0.x; s[1] = s0.y; s[2] = s0.z; s[3] = s0.w; s[4] = s1.x; s[5] = s1.y; s[6] = s1.z; s[7] = s1.w;
    }
    const uint4 raw = *reinterpret_cast<const uint4*>(p.xh + (size_t)row * kD + lane * 8);
    const float v[8] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y),
                        bf16lo(raw.z), bf16hi(raw.z), bf16lo(raw.w), bf16hi(raw.w)};
    float dsum = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) dsum = fmaf(v[k], s[k] - v[k], dsum);
    dsum = warp_sum(dsum);
    if (lane == 0) p.d[row] = dsum;
    acc += dsum;
  }
  if (lane == 0) { s_sum[warp] = acc; s_bag[warp] = cur; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int bag = -1;
    double t = 0.0;
    for (int w = 0; w < 8; ++w) {
      if (s_bag[w] < 0) continue;
      if (s_bag[w] != bag) {
        if (bag >= 0) atomicAdd(p.e + bag, t);
        bag = s_bag[w];
        t = 0.0;
      }
      t += (double)s_sum[w];
    }
    if (bag >= 0) atomicAdd(p.e + bag, t);
  }
}

// ------------------------------------------------------------------------------------------
// Degrees of a general (signed) patch graph:  d_i = sum_{j != i} relu(xh_i . xh_j),  e = sum_i d_i.
//   CTA = 128 rows x a range of 64-column tiles; A block resident in smem, B tiles by TMA (2 stages),
//   tcgen05 128x64x256 into two TMEM buffers; 8 epilogue warps, thread = (row, 32 of the 64 columns).
// ------------------------------------------------------------------------------------------
struct GramParams {
  const int* cu;
  const float* lfix;        // tiled, see lfix_index
  float* d;                 // (Rpad) degrees
  double* e;                // (B)
  float* T;                 // tiled like lfix, atomicAdd
  double* s;                // (B, 2 groups): sum_ij (A_ij/e - d_i d_j/e^2) delta_ij
  const int* negflag;       // (1) device flag: 1 when some element of h is negative (closed-form degrees do not apply)
  int tiles_per_split;      // column tiles per CTA
  float inv_temp;
  int row_lo, row_hi;       // rows owned by this call (sharded bag: this rank's row blocks); [0, INT_MAX) otherwise
};

constexpr size_t kDegSmem = 1024 + kABytes + kStages * kBBytes + 256;

__global__ void __launch_bounds__(kThreads, 1)
modularity_degrees_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                          const GramParams p) {
  if (!__ldg(p.negflag)) return;                      // all features >= 0: degrees came from the closed form
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
// Pair sweep (the trace and T): the dominant kernel of the training step.
//   CTA = 128 rows (bag-relative block) x a range of ABSOLUTE 64-column tiles; 16 warps, thread = (row i, 16 of
//   the 64 columns, four at a time), 128 registers per thread (a 17th warp would cap every thread at 96).
//   No dedicated producer / MMA warps: the pair work of a tile takes microseconds, so the sweep warps drive the
//   asynchronous machinery themselves, in rotation, with non-blocking mbarrier probes (five probe points per
//   tile): bulk copies of the L/degree tiles and the tcgen05 MMA 128x64x256 into a ring of six slots (TMEM
//   accumulator + L tile), up to five tiles ahead, and the TMA loads of the B boxes (two stages: with one, producing a
//   tile was the serial chain MMA retire -> probe -> TMA load -> probe -> MMA issue, ~4 us against 4.6 us of pair work
//   per tile, and paced the kernel: 32.8 -> 31.2 ms per 32 bags).
//   Per four columns a thread runs
//     - the (min,+) contraction over tokens: per token pair 2 broadcast 128-bit loads of L (token-major tile),
//       4 FADD2 (two columns each, row operand broadcast) and 4 FMNMX3 (one per column chain);
//     - the tanh / gradient tail on packed fp32x2 instructions, two columns per instruction;
//     - T_i[p*] += t as a fixed-point integer atomic on a token-major shared array (bank = row: conflict free).
// ------------------------------------------------------------------------------------------
constexpr int kSwWarps = 16;
constexpr int kSwThreads = kSwWarps * 32;           // 512
#ifdef IMP_SWEEP_TRACE
__device__ unsigned long long g_sweep_trace[16 * 8];   // debug counters (profiles/r01_sweep_iterations.md)
#endif
#ifndef IMP_SWEEP_BSTAGES
#define IMP_SWEEP_BSTAGES 2
#endif
constexpr int kBStages = IMP_SWEEP_BSTAGES;         // B boxes in flight: with one, the production of a tile is a serial chain
                                                    // MMA retire -> probe -> TMA load -> probe -> MMA issue of ~4 us against 4.6 us of pair work
constexpr int kTBufs = kBStages == 1 ? 8 : 6;       // ring slots: TMEM accumulator of 64 columns + L/degree tile.  One B stage: 4 slots 33.4 ms,
                                                    // 8 slots 32.9; two B stages: 6 slots 31.24, 7 slots 31.33 (and 1-4 probe points per tile within 0.4 ms)
constexpr int kLStages = kTBufs;
constexpr int kSwTmemCols = kTBufs * kBN <= 256 ? 256 : 512;   // allocations are powers of two

template <int NQ1, int NQ2>
constexpr size_t sweep_smem() {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  return 1024 + kABytes + kBStages * kBBytes + (size_t)kLStages * (kBN * PtPad * 4 + kBN * 4) + (size_t)PtPad * kBM * 4 + kBM * 4 + 64 + 512;
}

template <int NQ1, int NQ2, bool LASTPAD>      // LASTPAD: the last token slot is padding (never wins): skipped
__global__ void __launch_bounds__(kSwThreads, 1)
modularity_sweep_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                        const GramParams p) {
  constexpr int PtPad = 4 * (NQ1 + NQ2);
  constexpr int NG = NQ2 ? 2 : 1;
  constexpr int kLBytes = kBN * PtPad * 4;
  constexpr int kLStage = kLBytes + kBN * 4;            // L tile [PtPad][64] + degree tile [64]
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_a = smem;
  uint8_t* s_b = s_a + kABytes;
  uint8_t* s_l = s_b + kBStages * kBBytes;
  int* s_T = reinterpret_cast<int*>(s_l + kLStages * kLStage);           // [PtPad][128] fixed point, one row per patch
  float* s_invS = reinterpret_cast<float*>(s_T + (size_t)PtPad * kBM);   // [128] 1 / scale of the row
  float* s_dm = s_invS + kBM;                                            // [16] per-warp maxima of the degrees
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_dm + 16);
  uint64_t* bfull = bars;                 // [kBStages] TMA -> MMA
  uint64_t* bempty = bars + kBStages;     // [kBStages] MMA commit -> TMA
  uint64_t* full = bars + 2 * kBStages;   // [kTBufs] ring slot filled: L/degree bytes landed AND the MMAs retired
  uint64_t* empty = full + kTBufs;        // [kTBufs] ring slot consumed by all 16 warps
  uint64_t* afull = empty + kTBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(afull + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);                 // [16][2]
  volatile int* s_mcnt = reinterpret_cast<volatile int*>(s_red + 2 * kSwWarps);
  volatile uint64_t* s_desc = reinterpret_cast<volatile uint64_t*>(s_red + 2 * kSwWarps + 2);   // UMMA descriptors of A and B

  const int b = blockIdx.z;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int own_end = min(row_end, p.row_hi);                           // rows of this call: [max(row_begin,row_lo), own_end)
  const int i0 = max(row_begin, p.row_lo) + blockIdx.x * kBM;
  if (i0 >= own_end) return;
  const int ta0 = row_begin >> 6, ta1 = (row_end + kBN - 1) >> 6;        // absolute column tiles of the bag
  const int t0 = ta0 + blockIdx.y * p.tiles_per_split;
  const int ntiles = max(0, min(ta1, t0 + p.tiles_per_split) - t0);
  if (ntiles == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int i = 0; i < kBStages; ++i) { mbar_init(&bfull[i], 1); mbar_init(&bempty[i], 1); }
    for (int i = 0; i < kTBufs; ++i) { mbar_init(&full[i], 2); mbar_init(&empty[i], kSwWarps); }   // 2 = expect_tx arrive + commit
    mbar_init(afull, 1);
    *s_mcnt = 0;
    s_desc[0] = umma_desc_sw128(smem_u32(s_a), 0, 1024);
    s_desc[1] = umma_desc_sw128(smem_u32(s_b), 0, 1024);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kSwTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Producer work rotates over the 16 warps: warp n % 16 owns tile n and, from inside its own sweep, at non-blocking
  // probe points (five per tile), performs
  //   step 1 (gates: tiles < n issued, B box n landed, ring slot n % kTBufs released by all 16 warps):
  //           bulk copies of the L / degree tile n, the 16 tcgen05 MMAs of tile n, commits to bempty[n % 2] and full;
  //   step 2 (gate: the MMAs of tile n have retired = bempty[n % 2]): TMA load of B box n + 2 into the stage n frees.
  // Every step runs on a whole, converged warp and the asynchronous instructions on the lane elect.sync picks:
  // inside an elect-guarded block ptxas keeps descriptors and barrier addresses in uniform registers (~45
  // instructions for the 16 MMAs); behind a `lane == 0` test it wraps every UTCHMMA / UTMALDG in a per-lane
  // waterfall loop (~220).  Pinned roles (MMA on warp 13, B on 14, L on 15) made those warps the slowest of the
  // CTA - each mbarrier probe is a round trip through the busy shared-memory pipe, and the MMA issue stalls its
  // warp - and the other warps then sat at the accumulator barrier for 15% of the kernel (IMP_SWEEP_TRACE).
  // s_mcnt = number of tiles issued: a bfull barrier flips its phase every kBStages tiles, so it may only be probed
  // for tile n once tile n-1 has been issued (before that, its parity could still read as an older phase).
  int m_next = warp;                                   // next tile this warp owns (step 1 pending)
  int b_pend = -1;                                     // tile whose step 2 is pending, or -1
  auto issue_b = [&](int n) {                          // B box of tile n
    if (elect_one()) {
      const int bs = n % kBStages;
      mbar_arrive_expect_tx(&bfull[bs], kBBytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx)
        tma_load_2d(s_b + bs * kBBytes + bx * (kBN * 128), &tm_b, &bfull[bs], bx * 64, (t0 + n) * kBN);
    }
    __syncwarp();
  };
  auto step1 = [&]() {
    tc_fence_after();
    if (elect_one()) {
      const int slot = m_next % kTBufs;
      mbar_arrive_expect_tx(&full[slot], kLStage);
      uint8_t* dst = s_l + (size_t)slot * kLStage;
      const int ta = t0 + m_next;
      bulk_load(dst, p.lfix + (size_t)ta * PtPad * 64, kLBytes, &full[slot]);
      bulk_load(dst + kLBytes, p.d + (size_t)ta * 64, kBN * 4, &full[slot]);
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kBN, 0, 0);
      // base descriptors come from shared memory: computed here from the (loop-invariant) buffer addresses, ptxas
      // hoists the 32 additions below out of this block into every warp's tile loop (~40 instructions per tile)
      const uint64_t ad0 = s_desc[0], bd0 = s_desc[1] + (uint64_t)((m_next % kBStages) * (kBBytes >> 4));
      const uint32_t tacc = tmem_base + slot * kBN;
#pragma unroll
      for (int k = 0; k < kD / 16; ++k)      // descriptor start addresses advance in 16-byte units
        umma_f16(tacc, ad0 + (uint64_t)(((k >> 2) * (kBM * 128) + (k & 3) * 32) >> 4),
                 bd0 + (uint64_t)(((k >> 2) * (kBN * 128) + (k & 3) * 32) >> 4), idesc, k != 0);
      umma_commit(&bempty[m_next % kBStages]);
      umma_commit(&full[slot]);
      *s_mcnt = m_next + 1;
    }
    b_pend = (m_next + kBStages < ntiles) ? m_next : -1;     // the owner of tile n loads B box n + kBStages into the stage n frees
    m_next += kSwWarps;
    __syncwarp();
  };
  // non-blocking (every lane probes and the vote keeps the answer, and with it m_next / b_pend, warp-uniform)
  auto poll = [&](int it) {
    if (b_pend >= 0) {
      if (__all_sync(0xffffffffu, mbar_test(&bempty[b_pend % kBStages], (b_pend / kBStages) & 1))) { issue_b(b_pend + kBStages); b_pend = -1; }
    } else if (m_next < ntiles && m_next - it <= kTBufs) {      // its ring slot can be free at the earliest now
      const uint32_t par = ((m_next / kTBufs) & 1) ^ 1;
      const bool open = (*s_mcnt == m_next) && mbar_test(&bfull[m_next % kBStages], (m_next / kBStages) & 1) &&
                        mbar_test(&empty[m_next % kTBufs], par);
      if (__all_sync(0xffffffffu, open)) step1();
    }
  };
  // blocking: everything tile `it` needs from this warp has been issued.  The gates only depend on strictly
  // older tiles, which every warp that reached `it` has passed, and on steps their owners complete here.
  auto ensure = [&](int it) {
    if (b_pend >= 0 && b_pend < it) {
      mbar_wait_idle(&bempty[b_pend % kBStages], (b_pend / kBStages) & 1, 2000u);
      __syncwarp();
      issue_b(b_pend + kBStages);
      b_pend = -1;
    }
    if (m_next <= it) {
      const uint32_t par = ((m_next / kTBufs) & 1) ^ 1;
      while (*s_mcnt != m_next) {}
      mbar_wait_idle(&bfull[m_next % kBStages], (m_next / kBStages) & 1, 2000u);
      mbar_wait_idle(&empty[m_next % kTBufs], par, 2000u);
      __syncwarp();
      step1();
    }
  };
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(afull, kABytes);
#pragma unroll
      for (int bx = 0; bx < 4; ++bx) tma_load_2d(s_a + bx * (kBM * 128), &tm_a, afull, bx * 64, i0);
    }
    __syncwarp();
    for (int n = 0; n < kBStages && n < ntiles; ++n) issue_b(n);
  }
  mbar_wait(afull, 0);                                   // every warp issues MMAs that read the A tile

  const int q = warp & 3, hc = warp >> 2;
  const int tid = threadIdx.x;                         // = hc*128 + row in block
  const int i = i0 + q * 32 + lane;
  const bool row_ok = i < own_end;
  const uint32_t myT_s = smem_u32(s_T + q * 32 + lane);   // element p of this thread's row at byte offset p * 512
  float Li[PtPad];                                     // row operand: 2^23 + (N_i << 5)
#pragma unroll
  for (int k = 0; k < PtPad; ++k) {
    const float w = row_ok ? __ldg(p.lfix + lfix_index(i, k, PtPad)) : (float)(kNMax << 5);
    Li[k] = (float)((int)w & ~31) + 8388608.f;
  }
  for (int idx = tid; idx < PtPad * kBM; idx += kSwThreads) s_T[idx] = 0;
  const float di = row_ok ? __ldg(p.d + i) : 0.f;
  const double e = p.e[b];
  const float k1 = (float)(1.0 / e), k2 = (float)(1.0 / (e * e));
  const float gs4 = -800.f * p.inv_temp;                                 // 4 * 2 * (-100) / temp
  // T_i[p] is accumulated in 32-bit fixed point with native shared-memory integer atomics (ATOMS.ADD: 1.3 cycles
  // per conflict-free warp instruction; fp32 atomics on shared memory are CAS loops and the LDS/FADD/STS
  // read-modify-write it replaces serialised on the LDS latency), one array row per patch shared by the four
  // column-quarter warps.  Scale: a power of two S_i (exact in fp32) with sum_j |t_ij| S_i <= 2^30, from
  //   |t_ij| = |gs4 (A/e - d_i d_j/e^2)| * u e2/(1+e2)^2 <= (800/temp) max(A/e, d_i dmax/e^2) * 0.112 temp
  // (u/(4 cosh^2(u/temp)) peaks at u = 0.77 temp), A <= 1.02 for unit-norm bf16 rows.  Integer sums are
  // order-independent, so T no longer depends on the order in which warps and tiles are processed.
  float dmax = 0.f;
  for (int r = row_begin + tid; r < row_end; r += kSwThreads) dmax = fmaxf(dmax, __ldg(p.d + r));
  dmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(dmax)));      // degrees are >= 0
  if (lane == 0) s_dm[warp] = dmax;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kSwWarps; ++w) dmax = fmaxf(dmax, s_dm[w]);
  // at least 512 columns in the denominator: a single term stays below 2^21, inside the range of the magic-number rounding
  float sraw = 1073741824.f / ((float)max(ntiles * kBN, 512) * 92.f * fmaxf(k1, k2 * di * dmax));
  if (!(sraw < 1e30f)) sraw = 1e30f;
  if (!(sraw > 1e-30f)) sraw = 1e-30f;
  const float S = __int_as_float(__float_as_int(sraw) & 0x7f800000);     // 2^floor(log2 sraw)
  if (hc == 0) s_invS[q * 32 + lane] = 1.f / S;
  const float2 c0 = make_float2(gs4 * k1 * S, gs4 * k1 * S);
  const float2 nci = make_float2(-gs4 * k2 * di * S, -gs4 * k2 * di * S);
  const float2 magic2 = make_float2(12582912.f, 12582912.f);            // 1.5 * 2^23: (x + magic) has round(x) in its low bits
  const float nx2 = -2.f * 1.4426950408889634f * p.inv_temp;             // tanh(u/temp) via exp2(-2 u log2e / temp)
  const float2 nx2s = make_float2(nx2, nx2);
  const float2 kscale = make_float2(-1.f / (float)(1 << (kLogShift + 5)), -1.f / (float)(1 << (kLogShift + 5)));
  const float2 koff = make_float2(kArgOff, kArgOff);
  const float2 one2 = make_float2(1.f, 1.f), neg2 = make_float2(-1.f, -1.f);
  float2 sg[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};

#ifdef IMP_SWEEP_TRACE
  long long tr_t = 0, tr_l = 0, tr_e = 0, tr_look = 0;
  const long long tr_start = clock64();
#endif
  for (int it = 0; it < ntiles; ++it) {
    const int tb = it % kTBufs, ls = it % kLStages;
#ifdef IMP_SWEEP_TRACE
    const long long c0t = clock64();
#endif
    ensure(it);
    poll(it);
    __syncwarp();
#ifdef IMP_SWEEP_TRACE
    const long long c1t = clock64();
    tr_look += *s_mcnt - it;
#endif
    // sleeping waits: a warp that ran ahead must not spin away the issue slots of the warps it is waiting for
    // (try_wait with a suspend hint still came back every ~6 cycles here: 8% of all issued instructions)
    mbar_wait_sleep(&full[tb], (it / kTBufs) & 1);
    tc_fence_after();
#ifdef IMP_SWEEP_TRACE
    const long long c2t = clock64();
#endif
#ifdef IMP_SWEEP_TRACE
    const long long c3t = clock64();
    tr_e += c1t - c0t; tr_t += c2t - c1t; tr_l += c3t - c2t;
#endif
    const uint8_t* st = s_l + (size_t)ls * kLStage;
    const float4* sL = reinterpret_cast<const float4*>(st) + hc * 4;               // [token][16 float4]: + token*16 + g
    const float4* sd = reinterpret_cast<const float4*>(st + kLBytes) + hc * 4;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + tb * kBN + hc * 16;
    const int jt0 = (t0 + it) * kBN;
    const bool interior = (i0 + kBM <= own_end) && (jt0 >= row_begin) && (jt0 + kBN <= row_end) &&
                          (jt0 >= i0 + kBM || jt0 + kBN <= i0);

    auto group4 = [&](int g, auto interior_tag) {
      constexpr bool INTERIOR = decltype(interior_tag)::value;
      uint32_t v[4];
      tmem_ld4(taddr + g * 4, v);                       // Gram values of the 4 columns (consumed after the contraction)
      float m[4][2];
#pragma unroll
      for (int c = 0; c < 4; ++c) { m[c][0] = 3e38f; m[c][1] = 3e38f; }
#pragma unroll
      for (int pp = 0; pp < PtPad / 2; ++pp) {
        constexpr int dummy = 0; (void)dummy;
        const int grp = (2 * pp >= 4 * NQ1) ? 1 : 0;
        const float4 w0 = sL[(2 * pp) * 16 + g];
        const float2 la = make_float2(Li[2 * pp], Li[2 * pp]);
        const float2 a01 = add2(make_float2(w0.x, w0.y), la), a23 = add2(make_float2(w0.z, w0.w), la);
        if (LASTPAD && pp == PtPad / 2 - 1) {             // slot PtPad-1 holds the padding value for every row and column
          m[0][grp] = fminf(a01.x, m[0][grp]);
          m[1][grp] = fminf(a01.y, m[1][grp]);
          m[2][grp] = fminf(a23.x, m[2][grp]);
          m[3][grp] = fminf(a23.y, m[3][grp]);
          continue;
        }
        const float4 w1 = sL[(2 * pp + 1) * 16 + g];
        const float2 lb = make_float2(Li[2 * pp + 1], Li[2 * pp + 1]);
        const float2 b01 = add2(make_float2(w1.x, w1.y), lb), b23 = add2(make_float2(w1.z, w1.w), lb);
        m[0][grp] = fminf(fminf(a01.x, b01.x), m[0][grp]);
        m[1][grp] = fminf(fminf(a01.y, b01.y), m[1][grp]);
        m[2][grp] = fminf(fminf(a23.x, b23.x), m[2][grp]);
        m[3][grp] = fminf(fminf(a23.y, b23.y), m[3][grp]);
      }
      tmem_ld_wait();
      const float4 dj4 = sd[g];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {                  // column pairs (0,1) and (2,3)
        float2 a2 = make_float2(fmaxf(__uint_as_float(v[2 * h2]), 0.f), fmaxf(__uint_as_float(v[2 * h2 + 1]), 0.f));
        float2 dj2 = h2 ? make_float2(dj4.z, dj4.w) : make_float2(dj4.x, dj4.y);
        if (!INTERIOR) {
          const int j = jt0 + hc * 16 + g * 4 + 2 * h2;
          const bool ok0 = row_ok && j >= row_begin && j < row_end, ok1 = row_ok && j + 1 >= row_begin && j + 1 < row_end;
          a2.x = (ok0 && j != i) ? a2.x : 0.f;
          a2.y = (ok1 && j + 1 != i) ? a2.y : 0.f;
          dj2.x = ok0 ? dj2.x : 0.f;
          dj2.y = ok1 ? dj2.y : 0.f;
        }
        const float2 gw = fma2(a2, c0, mul2(dj2, nci));                 // 4*2*(-100)/temp * (A/e - d_i d_j/e^2); 0 when masked
#pragma unroll
        for (int grp = 0; grp < NG; ++grp) {
          const int m0 = __float_as_int(m[2 * h2][grp]), m1 = __float_as_int(m[2 * h2 + 1][grp]);
          const int pl0 = m0 & 31, pl1 = m1 & 31;
          const float2 F = make_float2(__int_as_float(m0 & ~31), __int_as_float(m1 & ~31));
          const float2 arg = fma2(F, kscale, koff);                     // log2(C_i C_j) of the winning token
          const float2 u = make_float2(ex2_approx(arg.x), ex2_approx(arg.y));
          const float2 x = mul2(u, nx2s);
          const float2 e2 = make_float2(ex2_approx(x.x), ex2_approx(x.y));   // exp(-2u/temp)
          const float2 den = add2(e2, one2);
          const float2 r = make_float2(rcp_approx(den.x), rcp_approx(den.y));
          const float2 t1 = mul2(e2, r);
          const float2 delta = fma2(t1, neg2, r);                       // tanh(u/temp) = (1-e2)/(1+e2)
          sg[grp] = fma2(gw, delta, sg[grp]);                           // sum of S gs4 * (A/e - d_i d_j/e^2) * delta
          // t = 2 g (1-delta^2)/temp * u, 1-delta^2 = 4 e2 r^2, times S, rounded to an integer by the magic addend
          const float2 tf = fma2(mul2(gw, t1), mul2(r, u), magic2);
          red_shared_add(myT_s + (grp ? 4 * NQ1 * kBM * 4 : 0), pl0, __float_as_int(tf.x) - 0x4B400000);
          red_shared_add(myT_s + (grp ? 4 * NQ1 * kBM * 4 : 0), pl1, __float_as_int(tf.y) - 0x4B400000);
        }
      }
    };
    auto tile = [&](auto interior_tag) {
#pragma unroll 2      // not 4: the fully unrolled tile (21 KB of SASS per variant) stalled on instruction fetch (37.6 -> 35.8 ms); 1 is slower (38.6)
      for (int g = 0; g < 4; ++g) {       // probe points per tile, measured with ONE B stage (ms per 32 bags): 1: 35.9 | 2: 33.2 | 4: 32.8 | 6: 33.8 | 8: 34.8
        group4(g, interior_tag);
        poll(it);
        __syncwarp();
      }
    };
    if (interior) tile(std::true_type{}); else tile(std::false_type{});
    tc_fence_before();                                   // accumulator and L tile of the slot have been read completely
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[tb]);
  }
#ifdef IMP_SWEEP_TRACE
  if (lane == 0) {     // per warp: cycles waiting for the accumulator / the L tile / inside ensure(), loop cycles, tiles, look-ahead
    atomicAdd(&g_sweep_trace[warp * 8 + 0], (unsigned long long)tr_t);
    atomicAdd(&g_sweep_trace[warp * 8 + 1], (unsigned long long)tr_l);
    atomicAdd(&g_sweep_trace[warp * 8 + 2], (unsigned long long)(clock64() - tr_start));
    atomicAdd(&g_sweep_trace[warp * 8 + 3], (unsigned long long)ntiles);
    atomicAdd(&g_sweep_trace[warp * 8 + 4], (unsigned long long)tr_look);
    atomicAdd(&g_sweep_trace[warp * 8 + 5], (unsigned long long)tr_e);
  }
#endif
  // ---------------- flush ----------------
  {
    const float inv_gs4 = s_invS[q * 32 + lane] / gs4;
#pragma unroll
    for (int grp = 0; grp < 2; ++grp) {
      const float a1 = warp_sum((sg[grp].x + sg[grp].y) * inv_gs4);   // = sum A delta / e - sum d_i d_j delta / e^2
      if (lane == 0) s_red[warp * 2 + grp] = a1;
    }
  }
  __syncthreads();
  if (tid < 2) {
    double t = 0.0;
    for (int w = 0; w < kSwWarps; ++w) t += (double)s_red[w * 2 + tid];
    atomicAdd(p.s + (size_t)b * 2 + tid, t);
  }
  // one RED per (row, token)
  for (int idx = tid; idx < kBM * PtPad; idx += kSwThreads) {
    const int r = idx & (kBM - 1), k = idx >> 7;
    const int v = s_T[idx];
    if (v != 0 && i0 + r < own_end) atomicAdd(p.T + lfix_index(i0 + r, k, PtPad), (float)v * s_invS[r]);   // tiled like L: coalesced REDs
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kSwTmemCols);
  }
}

// ------------------------------------------------------------------------------------------
// finish: dC = T / C (C = exp2(kCOff - N/2^kLogShift)), dchat_p = sum_i dC_ip xh_i ; loss per group
// ------------------------------------------------------------------------------------------
struct FinishParams {
  const bf16* h;
  const float* invn;
  const float* lfix;
  const float* T;
  const int* cu;
  const double* s;
  float* dchat;       // (B, Pt, 256), zeroed by the launcher
  float* loss;        // (B, 2)
  int P1, P2, P1pad, rows_per_cta;    // rows_per_cta: 64-row tiles per CTA
  int row_lo, row_hi;       // rows owned by this call; h points at global row h_row0
  int h_row0;
};

// ------------------------------------------------------------------------------------------
// finish on tcgen05:  dchat[slot][f] = sum_rows (dC[row][slot] / |x_row|) * h[row][f]   (a PtPad x N x 256 GEMM per bag)
//   per absolute 64-row tile: TMA box of h (exact in bf16: the MN-major B operand), bulk copies of the T and L tiles
//   (token-major); four builder warps turn them into a = T / C / |x| split into two bf16 MN-major A operands
//   [64 rows][64 slots] (a = hi + lo keeps ~16 mantissa bits), one thread issues 2 x 4 MMAs 128 x 256 x 16 into a TMEM
//   accumulator that lives for the whole CTA (rows 64..127 of A are a zero box); at the end the builder warps read
//   the PtPad real lanes and add them to dchat.
// ------------------------------------------------------------------------------------------
constexpr int kFtStages = 3;
constexpr int kFtThreads = 6 * 32;
template <int PTPAD>
constexpr size_t finish_tc_smem() { return 1024 + (size_t)kFtStages * (kBBytes + 2 * PTPAD * 256 + 16384) + 8192 + 256; }

template <int PTPAD>
__global__ void __launch_bounds__(kFtThreads, 1)
modularity_finish_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const FinishParams p) {
  constexpr int kTL = PTPAD * 256;                       // bytes of one T (or L) tile: [PTPAD][64] fp32
  constexpr int kStage = kBBytes + 2 * kTL + 16384;      // h box | T tile | L tile | A hi box | A lo box
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* s_zero = smem + kFtStages * kStage;           // second 64-wide M box of A: zeros
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_zero + 8192);
  uint64_t* full = bars;                  // [3] loads -> builders + MMA
  uint64_t* afull = bars + kFtStages;     // [3] 4 builder warps -> MMA
  uint64_t* empty = afull + kFtStages;    // [3] MMA commit -> producer
  uint64_t* tfull = empty + kFtStages;    // all MMAs retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int b = blockIdx.y;
  const int row_begin = __ldg(p.cu + b), row_end = __ldg(p.cu + b + 1);
  const int Pt = p.P1 + p.P2;
  if (blockIdx.x == 0 && threadIdx.x < 2)
    p.loss[b * 2 + threadIdx.x] = (float)(-100.0 * p.s[(size_t)b * 2 + threadIdx.x]);     // utils.py:222-228
  const int lo = max(row_begin, p.row_lo), hi = min(row_end, p.row_hi);
  if (lo >= hi) return;
  const int ta0 = lo >> 6, ta1 = (hi + 63) >> 6;
  const int t0 = ta0 + blockIdx.x * p.rows_per_cta;      // rows_per_cta: here tiles per CTA
  const int ntiles = max(0, min(ta1, t0 + p.rows_per_cta) - t0);
  if (ntiles == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < kFtStages; ++i) { mbar_init(&full[i], 1); mbar_init(&afull[i], 4); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  // A boxes and the zero box start as zeros: builders only ever write the PTPAD/8 real 16-byte chunks of a row
  for (int i = threadIdx.x; i < (kFtStages * 16384 + 8192) / 16; i += kFtThreads) {
    const int st = i / 1024, o = i % 1024;
    uint8_t* dst = st < kFtStages ? smem + (size_t)st * kStage + kBBytes + 2 * kTL : s_zero;
    reinterpret_cast<uint4*>(dst)[o] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      for (int it = 0; it < ntiles; ++it) {
        const int st = it % kFtStages;
        mbar_wait_idle(&empty[st], ((it / kFtStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], kBBytes + 2 * kTL);
        uint8_t* dst = smem + (size_t)st * kStage;
        const int ta = t0 + it;
#pragma unroll
        for (int bx = 0; bx < 4; ++bx) tma_load_2d(dst + bx * (kBN * 128), &tm_x, &full[st], bx * 64, ta * 64 - p.h_row0);
        bulk_load(dst + kBBytes, p.T + (size_t)ta * PTPAD * 64, kTL, &full[st]);
        bulk_load(dst + kBBytes + kTL, p.lfix + (size_t)ta * PTPAD * 64, kTL, &full[st]);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);
      for (int it = 0; it < ntiles; ++it) {
        const int st = it % kFtStages;
        mbar_wait_idle(&afull[st], (it / kFtStages) & 1);      // implies full[st]: the builders waited for it
        tc_fence_after();
        const uint32_t sb = smem_u32(smem + (size_t)st * kStage);
        const uint32_t sa = sb + kBBytes + 2 * kTL;
#pragma unroll
        for (int half = 0; half < 2; ++half) {                 // hi and lo parts of the A operand
          const uint32_t sah = sa + half * 8192;
          const uint32_t lbo_a = smem_u32(s_zero) - sah;       // second M box = the shared zero box
#pragma unroll
          for (int k = 0; k < 4; ++k) {                        // 16 rows per MMA; MN-major: +2048 B per 16 k-rows
            const uint64_t ad = umma_desc_sw128(sah + k * 2048, lbo_a, 1024);
            const uint64_t bd = umma_desc_sw128(sb + k * 2048, kBN * 128, 1024);
            umma_f16(tmem_base, ad, bd, idesc, (it | half | k) != 0);
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(tfull);
    }
  } else {
    // ---------------- builders: dC tile as a bf16 MN-major A operand ----------------
    const int tid = threadIdx.x;                               // 0..127
    for (int it = 0; it < ntiles; ++it) {
      const int st = it % kFtStages;
      mbar_wait(&full[st], (it / kFtStages) & 1);
      const uint8_t* base = smem + (size_t)st * kStage + kBBytes;
      const float* sT = reinterpret_cast<const float*>(base);
      const float* sL = reinterpret_cast<const float*>(base + kTL);
      uint8_t* sA = const_cast<uint8_t*>(base) + 2 * kTL;
      const int row0 = (t0 + it) * 64;
      for (int item = tid; item < 64 * (PTPAD / 8); item += 128) {
        const int r = item & 63, ch = item >> 6;
        const bool ok = row0 + r >= lo && row0 + r < hi;
        const float inv = ok ? __ldg(p.invn + row0 + r) : 0.f;
        uint32_t w[4], wl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v[2], vl[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int slot = ch * 8 + 2 * e + h2;
            const float t = sT[slot * 64 + r];
            const int nfix = (int)sL[slot * 64 + r] >> 5;
            const float c = ex2_approx((float)kCOff - (float)nfix * (1.f / (float)(1 << kLogShift)));
            v[h2] = (ok && t != 0.f && nfix < kNMax) ? t / c * inv : 0.f;   // relu gate: C == 0 never wins the max with u > 0
            vl[h2] = v[h2] - __bfloat162float(__float2bfloat16_rn(v[h2]));
          }
          w[e] = pack_bf16x2(v[0], v[1]);
          wl[e] = pack_bf16x2(vl[0], vl[1]);
        }
        const uint32_t off = r * 128 + ((ch ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(sA + off) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(sA + 8192 + off) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[st]);
    }
    // ---------------- epilogue: lanes = token slots ----------------
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int slot = warp * 32 + lane;
    int tok = -1;
    if (slot < p.P1) tok = slot;
    else if (slot >= p.P1pad && slot - p.P1pad < p.P2) tok = p.P1 + slot - p.P1pad;
    if (warp * 32 < PTPAD) {                                   // warps whose lanes hold real slots
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cc * 32, v);
        tmem_ld_wait();
        if (tok >= 0) {
          float* dst = p.dchat + ((size_t)b * Pt + tok) * kD + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = __uint_as_float(v[j]);
            if (a != 0.f) atomicAdd(dst + j, a);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// Column split of the pair sweep: grid = (row blocks of 128, nsplit, bags), a CTA covers tiles_per_split 64-column tiles.
//  - at least 16 column tiles per CTA (the A block load, the row operand and the flush are per CTA: ~0.5 tile);
//  - at most `max_tiles` (256 = 16 384 columns): a CTA accumulates T_i over its own columns in 32-bit fixed point whose
//    scale is set by the column count, which keeps the resolution of a 120 000-patch bag at that of a 16 384-patch one;
//  - among the admissible counts (up to four times the smallest) the one with the fewest wave-steps
//    ceil(CTAs / SMs) x (tiles_per_split + 0.5): a rank's share of a sharded giant bag is ~118 row blocks, and eight
//    splits of 235 tiles were 6.4 waves = seven steps of 235 where ten splits of 188 are eight full steps of 188 (-8 %).
void modularity_sweep_plan(int own_len, int max_len, int B, int* nsplit_out, int* tiles_per_split_out) {
  const int sms = imp_num_sms();
  const int row_blocks = (own_len + kBM - 1) / kBM;
  const int col_tiles = (max_len + kBN - 1) / kBN + 1;        // absolute tiles: a bag may straddle one more
  static const int max_tiles = []() { const char* e = getenv("IMP_SWEEP_MAXTILES"); return e && atoi(e) > 0 ? atoi(e) : 256; }();
  const int ns_cap = std::max(1, col_tiles / 16);
  int ns_min = std::max(1, std::min((2 * sms + row_blocks * B - 1) / (row_blocks * B), ns_cap));     // >= 2 waves
  ns_min = std::max(ns_min, (col_tiles + max_tiles - 1) / max_tiles);
  const int ns_max = std::max(ns_min, std::min(4 * ns_min, ns_cap));
  long best_cost = -1;
  int best_ns = ns_min, best_tps = (col_tiles + ns_min - 1) / ns_min;
  for (int ns = ns_min; ns <= ns_max; ++ns) {
    const int tps = (col_tiles + ns - 1) / ns;
    const int ns_eff = (col_tiles + tps - 1) / tps;
    const long ctas = (long)row_blocks * B * ns_eff;
    const long cost = ((ctas + sms - 1) / sms) * (2L * tps + 1);
    // a different split must buy at least 2 % (the wave model ignores that CTAs of a partial last wave overlap the previous one)
    if (best_cost < 0 || cost * 100 < best_cost * 98) { best_cost = cost; best_ns = ns_eff; best_tps = tps; }
  }
  *nsplit_out = best_ns;
  *tiles_per_split_out = best_tps;
}

namespace {

int quads1(int P1) { return P1 <= 8 ? 2 : (P1 <= 16 ? 4 : 8); }
int quads2(int P2) { return P2 == 0 ? 0 : 2; }

int run_degrees(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  { const int rc_ = imp_ensure_smem((const void*)modularity_degrees_kernel, kDegSmem); if (rc_) return rc_; }
  IMP_LAUNCH("modularity_degrees_gram", st, modularity_degrees_kernel<<<grid, kThreads, kDegSmem, st>>>(ta, tb, p));
  return IMP_OK;
}

template <int NQ1, int NQ2, bool LASTPAD = false>
int run_sweep(const CUtensorMap& ta, const CUtensorMap& tb, const GramParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t smem = sweep_smem<NQ1, NQ2>();
  static_assert(smem <= 227 * 1024, "modularity_sweep shared memory");
  { const int rc_ = imp_ensure_smem((const void*)modularity_sweep_kernel<NQ1, NQ2, LASTPAD>, smem); if (rc_) return rc_; }
  IMP_LAUNCH("modularity_sweep", st, modularity_sweep_kernel<NQ1, NQ2, LASTPAD><<<grid, kSwThreads, smem, st>>>(ta, tb, p));
  return IMP_OK;
}

template <int PTPAD>
int run_prep_tc(const PrepParams& pp, cudaStream_t st) {
  constexpr size_t smem = prep_tc_smem<PTPAD>();
  CUtensorMap tm;
  int rc = imp_make_tmap_2d(&tm, pp.h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, pp.row_hi - pp.row_lo, kD * 2, 64, kBM,
                            CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  { const int rc_ = imp_ensure_smem((const void*)modularity_prep_tc_kernel<PTPAD>, smem); if (rc_) return rc_; }
  IMP_LAUNCH("modularity_prep", st, modularity_prep_tc_kernel<PTPAD><<<(pp.row_hi - pp.row_lo + kBM - 1) / kBM, kPtThreads, smem, st>>>(tm, pp));
  return IMP_OK;
}

template <int PTPAD>
int run_prep(const PrepParams& pp, cudaStream_t st) { return run_prep_tc<PTPAD>(pp, st); }

struct Carve {
  bf16* xh; float* invn; float* lfix; float* d; float* T; double* e; double* s; float* colsum; int* negflag;
  size_t zero_off, zero_bytes, total;
};
Carve carve(void* ws, int total_rows, int B, int PtPad) {
  const size_t rpad = (size_t)total_rows + kBM + kBN;
  const size_t ntile = (rpad + 63) / 64;
  auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  size_t off = 0;
  Carve c;
  c.xh = reinterpret_cast<bf16*>(base + off); off += up(rpad * kD * 2);
  c.invn = reinterpret_cast<float*>(base + off); off += up(rpad * 4);
  c.lfix = reinterpret_cast<float*>(base + off); off += up(ntile * PtPad * 64 * 4);
  c.zero_off = off;
  c.d = reinterpret_cast<float*>(base + off); off += up(rpad * 4);
  c.T = reinterpret_cast<float*>(base + off); off += up(ntile * PtPad * 64 * 4);
  c.e = reinterpret_cast<double*>(base + off); off += up((size_t)B * 8);
  c.s = reinterpret_cast<double*>(base + off); off += up((size_t)B * 2 * 8);
  c.colsum = reinterpret_cast<float*>(base + off); off += up((size_t)B * kD * 4);
  c.negflag = reinterpret_cast<int*>(base + off); off += up(4);
  c.zero_bytes = off - c.zero_off;
  c.total = off;
  return c;
}

}  // namespace

size_t modularity_workspace_bytes(int total_rows, int B, int P1, int P2) {
  const int PtPad = 4 * (quads1(P1) + quads2(P2));
  return carve(nullptr, total_rows, B, PtPad).total + 256;
}

namespace {
int check_modularity_args(int total_rows, int B, int max_len, int P1, int P2, float temp, const void* workspace) {
  if (P1 < 1 || P1 > 32 || P2 < 0 || P2 > 8) IMP_FAIL(IMP_ERR_ARG, "modularity: token groups (%d,%d) must be in [1,32] and [0,8]", P1, P2);
  if (total_rows <= 0 || max_len <= 0) IMP_FAIL(IMP_ERR_ARG, "modularity: empty input");
  if (!(temp > 0.f)) IMP_FAIL(IMP_ERR_ARG, "modularity: temp must be positive");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) IMP_FAIL(IMP_ERR_ARG, "modularity: workspace must be 256-byte aligned");
  (void)B;
  return IMP_OK;
}
}  // namespace

// sections of the workspace a multi-GPU caller exchanges between the two phases (byte offsets and sizes):
//   0 xh (row-major, 512 B per row)   1 lfix (tiled, PtPad*256 B per 64-row tile)   2 colsum (B*256 floats)   3 negflag (1 int)
void modularity_workspace_sections(int total_rows, int B, int P1, int P2, size_t* offsets, size_t* sizes) {
  const int PtPad = 4 * (quads1(P1) + quads2(P2));
  Carve c = carve(nullptr, total_rows, B, PtPad);
  const uint8_t* base = nullptr;
  const size_t ntile = ((size_t)total_rows + kBM + kBN + 63) / 64;
  offsets[0] = reinterpret_cast<const uint8_t*>(c.xh) - base;      sizes[0] = (size_t)total_rows * kD * 2;
  offsets[1] = reinterpret_cast<const uint8_t*>(c.lfix) - base;    sizes[1] = ntile * PtPad * 64 * 4;
  offsets[2] = reinterpret_cast<const uint8_t*>(c.colsum) - base;  sizes[2] = (size_t)B * kD * 4;
  offsets[3] = reinterpret_cast<const uint8_t*>(c.negflag) - base; sizes[3] = 4;
}

// Phase 1: zero the accumulators and run the prep kernel over the global rows [row_lo, row_lo + local_rows) whose
// features are at h_local (a whole batch: row_lo = 0, local_rows = total_rows).
int launch_modularity_prepare(const bf16* h_local, int local_rows, int row_lo, int total_rows, const int* cu, int B,
                              const float* chat, int P1, int P2, void* workspace, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  int rc = check_modularity_args(total_rows, B, 1, P1, P2, 1.f, workspace);
  if (rc) return rc;
  if (row_lo < 0 || local_rows < 0 || row_lo + local_rows > total_rows || (row_lo & 63))
    IMP_FAIL(IMP_ERR_ARG, "modularity: row window [%d,+%d) must lie in [0,%d) and start on a multiple of 64", row_lo, local_rows, total_rows);
  const int nq1 = quads1(P1), nq2 = quads2(P2);
  const int P1pad = 4 * nq1, PtPad = 4 * (nq1 + nq2);
  Carve c = carve(workspace, total_rows, B, PtPad);
  // padding rows of xh are read by the tile loads of the last row block / column tile: keep them finite
  IMP_CUDA(cudaMemsetAsync(c.xh + (size_t)total_rows * kD, 0, (size_t)(kBM + kBN) * kD * 2, st));
  IMP_CUDA(cudaMemsetAsync(reinterpret_cast<uint8_t*>(workspace) + c.zero_off, 0, c.zero_bytes, st));
  if (local_rows == 0) return IMP_OK;
  PrepParams pp;
  pp.h = h_local; pp.cu = cu; pp.chat = chat; pp.xh = c.xh; pp.invn = c.invn; pp.lfix = c.lfix; pp.colsum = c.colsum;
  pp.negflag = c.negflag; pp.R = total_rows; pp.B = B; pp.P1 = P1; pp.P2 = P2; pp.P1pad = P1pad;
  pp.row_lo = row_lo; pp.row_hi = row_lo + local_rows;
  switch (PtPad) {
    case 8: rc = run_prep<8>(pp, st); break;
    case 16: rc = run_prep<16>(pp, st); break;
    case 24: rc = run_prep<24>(pp, st); break;
    case 32: rc = run_prep<32>(pp, st); break;
    case 40: rc = run_prep<40>(pp, st); break;
    default: IMP_FAIL(IMP_ERR_ARG, "modularity: unsupported padded token count %d", PtPad);
  }
  return rc;
}

// Phase 2: degrees over all rows, then the pair sweep and the finish kernel for the rows [row_lo, row_lo + local_rows)
// against all columns.  For a window smaller than the batch, `loss` and `dchat` are this window's partial sums.
int launch_modularity_execute(const bf16* h_local, int local_rows, int row_lo, int total_rows, const int* cu, int B,
                              int max_len, int P1, int P2, float temp, void* workspace, float* loss, float* dchat,
                              cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  int rc = check_modularity_args(total_rows, B, max_len, P1, P2, temp, workspace);
  if (rc) return rc;
  const int nq1 = quads1(P1), nq2 = quads2(P2);
  const int P1pad = 4 * nq1, PtPad = 4 * (nq1 + nq2), Pt = P1 + P2;
  Carve c = carve(workspace, total_rows, B, PtPad);
  IMP_CUDA(cudaMemsetAsync(dchat, 0, (size_t)B * Pt * kD * 4, st));
  const bool whole = (row_lo == 0 && local_rows == total_rows);
  const int own_len = whole ? max_len : local_rows;          // longest run of owned rows inside one bag

  DegParams dp;
  dp.xh = c.xh; dp.cu = cu; dp.colsum = c.colsum; dp.negflag = c.negflag; dp.d = c.d; dp.e = c.e; dp.R = total_rows; dp.B = B;
  IMP_LAUNCH("modularity_degrees_closed", st, modularity_degrees_closed_kernel<<<(total_rows + 63) / 64, 256, 0, st>>>(dp));

  CUtensorMap ta, tb;
  const uint64_t rpad = (uint64_t)total_rows + kBM + kBN;
  if ((rc = imp_make_tmap_2d(&ta, c.xh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rpad, kD * 2, 64, kBM, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = imp_make_tmap_2d(&tb, c.xh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rpad, kD * 2, 64, kBN, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  GramParams gp;
  gp.cu = cu; gp.lfix = c.lfix; gp.d = c.d; gp.e = c.e; gp.T = c.T; gp.s = c.s; gp.negflag = c.negflag; gp.inv_temp = 1.f / temp;
  gp.row_lo = 0; gp.row_hi = INT_MAX;
  {   // general (signed) degrees: every row of every bag, exits at once unless some feature is negative
    const int row_blocks = (max_len + kBM - 1) / kBM;
    const int col_tiles = (max_len + kBN - 1) / kBN;
    int nsplit = std::max(1, std::min((2 * imp_num_sms() + row_blocks * B - 1) / (row_blocks * B), std::max(1, col_tiles / 16)));
    gp.tiles_per_split = (col_tiles + nsplit - 1) / nsplit;
    nsplit = (col_tiles + gp.tiles_per_split - 1) / gp.tiles_per_split;
    if ((rc = run_degrees(ta, tb, gp, dim3(row_blocks, nsplit, B), st))) return rc;
  }
  if (local_rows > 0) {
    gp.row_lo = whole ? 0 : row_lo; gp.row_hi = whole ? INT_MAX : row_lo + local_rows;
    const int row_blocks = (own_len + kBM - 1) / kBM;
    const int col_tiles = (max_len + kBN - 1) / kBN + 1;        // absolute tiles: a bag may straddle one more
    int nsplit;
    modularity_sweep_plan(own_len, max_len, B, &nsplit, &gp.tiles_per_split);
    const dim3 grid(row_blocks, nsplit, B);
#define IMP_SWEEP(a, b2) rc = run_sweep<a, b2>(ta, tb, gp, grid, st)
    if (nq2 == 0) { if (nq1 == 2) IMP_SWEEP(2, 0); else if (nq1 == 4) IMP_SWEEP(4, 0); else IMP_SWEEP(8, 0); }
    else if (nq1 == 8 && (P2 & 3)) rc = run_sweep<8, 2, true>(ta, tb, gp, grid, st);     // the 32 + 7 token configuration
    else { if (nq1 == 2) IMP_SWEEP(2, 2); else if (nq1 == 4) IMP_SWEEP(4, 2); else IMP_SWEEP(8, 2); }
#undef IMP_SWEEP
    if (rc) return rc;
  }

  FinishParams fp;
  fp.h = h_local; fp.invn = c.invn; fp.lfix = c.lfix; fp.T = c.T; fp.cu = cu; fp.s = c.s;
  fp.dchat = dchat; fp.loss = loss; fp.P1 = P1; fp.P2 = P2; fp.P1pad = P1pad;
  fp.row_lo = whole ? 0 : row_lo; fp.row_hi = whole ? INT_MAX : row_lo + local_rows; fp.h_row0 = row_lo;
  if (local_rows == 0) {        // a rank without rows of the bag: its partial sums are zero
    IMP_CUDA(cudaMemsetAsync(loss, 0, (size_t)B * 2 * 4, st));
    return IMP_OK;
  }
  // tcgen05 finish: about one CTA per SM, each over a contiguous run of absolute 64-row tiles of one bag
  CUtensorMap th;          // the rows of h this call owns (out-of-range rows of a tile are zero-filled by TMA)
  if ((rc = imp_make_tmap_2d(&th, h_local, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, (uint64_t)std::max(local_rows, 1), kD * 2, 64, kBN,
                             CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const int fin_tiles = (std::max(own_len, 1) + 63) / 64 + 1;
  // one CTA per SM (218 KB of shared memory): chunks per bag that fill whole waves, cost = waves * (tiles per CTA + 2)
  // (rounding up to ceil(SMs / B) chunks gave 160 CTAs on 148 SMs for 32 bags: a second wave of 12)
  int fin_chunks = 1;
  {
    long best = -1;
    for (int ch = 1; ch <= std::min(fin_tiles, 2 * imp_num_sms()); ++ch) {
      const int tps = (fin_tiles + ch - 1) / ch, real = (fin_tiles + tps - 1) / tps;
      const long cost = (((long)B * real + imp_num_sms() - 1) / imp_num_sms()) * (tps + 2);
      if (best < 0 || cost < best) { best = cost; fin_chunks = real; }
    }
  }
  fp.rows_per_cta = (fin_tiles + fin_chunks - 1) / fin_chunks;           // tiles per CTA
  const dim3 fgrid((fin_tiles + fp.rows_per_cta - 1) / fp.rows_per_cta, B);
#define IMP_FIN(PT)                                                                                                    \
  do {                                                                                                                  \
    constexpr size_t smem = finish_tc_smem<PT>();                                                                       \
  { const int rc_ = imp_ensure_smem((const void*)modularity_finish_tc_kernel<PT>, smem); if (rc_) return rc_; }                                                                                                                   \
    IMP_LAUNCH("modularity_finish", st, modularity_finish_tc_kernel<PT><<<fgrid, kFtThreads, smem, st>>>(th, fp));     \
  } while (0)
  switch (PtPad) {
    case 8: IMP_FIN(8); break;
    case 16: IMP_FIN(16); break;
    case 24: IMP_FIN(24); break;
    case 32: IMP_FIN(32); break;
    case 40: IMP_FIN(40); break;
    default: IMP_FAIL(IMP_ERR_ARG, "modularity: unsupported padded token count %d", PtPad);
  }
#undef IMP_FIN
  IMP_LAUNCH_CHECK();
  return IMP_OK;
}

int launch_modularity(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* chat, int P1, int P2,
                      float temp, void* workspace, float* loss, float* dchat, cudaStream_t st) {
  if (B <= 0) return IMP_OK;
  int rc = check_modularity_args(total_rows, B, max_len, P1, P2, temp, workspace);
  if (rc) return rc;
  if ((rc = launch_modularity_prepare(h, total_rows, 0, total_rows, cu, B, chat, P1, P2, workspace, st))) return rc;
  return launch_modularity_execute(h, total_rows, 0, total_rows, cu, B, max_len, P1, P2, temp, workspace, loss, dchat, st);
}

#ifdef IMP_SWEEP_TRACE
extern "C" int imp_debug_sweep_trace(unsigned long long* host_out, int reset) {
  if (host_out) cudaMemcpyFromSymbol(host_out, g_sweep_trace, sizeof(unsigned long long) * 128);
  if (reset) { static unsigned long long z[128] = {0}; cudaMemcpyToSymbol(g_sweep_trace, z, sizeof(z)); }
  return 0;
}
#endif
