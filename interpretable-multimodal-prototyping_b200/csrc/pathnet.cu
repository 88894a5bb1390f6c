// A1  path_net on tcgen05 tensor cores (reference: medmm/modeling/models/umeml_gan.py:266-268,410)
//
//   forward : h[r, :]  = dropout(relu(x[r, :] W1^T + b1))          x (R,512) bf16, W1 (256,512) bf16
//   backward: dW1      = dz^T x   (256,512), split-K over the patch rows, fp32 partials
//
// Both are warp-specialised persistent kernels: one TMA producer warp, one MMA-issuer warp
// (a single thread issues tcgen05.mma, accumulators live in TMEM) and epilogue warps that read
// TMEM with tcgen05.ld.  Operands are staged in shared memory by TMA with the 128-byte swizzle.
#include "common.cuh"
#include <stdlib.h>

namespace {

constexpr int kD = 256;        // hidden dim (MODEL.HIDDEN_DIM)
constexpr int kBM = 128;       // patch rows per tile
constexpr int kBK = 64;        // K elements per stage (one 128 B swizzle row of bf16)
// Separate rings for x (HBM) and the W1 slabs every tile re-reads (L2 hits), depths as template parameters.  Measured on
// 32 bags x 16 384 patches (ms): (x,W) = (4,4) 0.169 | (5,3) 0.167 | (6,3) 0.172 | (8,2) 0.188.  Variants that were built,
// measured and dropped: W1 stage multicast to a CTA pair 0.165; 16 epilogue warps 0.184; h tile
// through shared memory + TMA store 0.186; one half of W1 resident in shared memory (x-only ring of 5 stages) 0.184.
// Ablation of this kernel: loads only 0.123, + MMAs 0.142, + epilogue 0.145, all 0.169-0.180: the ring (192 KB in
// flight per SM) cannot hide the load latency once slots are also held for the MMAs.
constexpr int kEpiWarps = 8;       // two per TMEM lane quarter, 128 output columns each
constexpr int kColsPerWarp = kD / (kEpiWarps / 4);
constexpr int kChunks = kColsPerWarp / 32;
constexpr int kFwdThreads = (kEpiWarps + 2) * 32;
constexpr uint32_t kStageBytesA = kBM * kBK * 2;     // 16 KB
constexpr uint32_t kStageBytesB = kD * kBK * 2;      // 32 KB
template <int kXStages, int kWStages>
constexpr size_t fwd_smem() { return 1024 + (size_t)kXStages * kStageBytesA + (size_t)kWStages * kStageBytesB + 256 * 4 + 256; }

struct FwdParams {
  const float* bias;      // (256)
  bf16* h;                // (R,256)
  int rows;
  int num_tiles;
  int kdim;               // 512 (multiple of 64)
  float keep_scale;       // 1/(1-p)
  uint32_t drop_thresh;   // keep element iff (16 hash bits) >= drop_thresh = round(65536 p); 0 => no dropout
  uint32_t seed;
  const uint32_t* seed_offset;   // device word XOR-ed into the seed, or null
};

template <int kXStages, int kWStages>
__global__ void __launch_bounds__(kFwdThreads, 1)
pathnet_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* x_base = smem;
  uint8_t* w_base = smem + (size_t)kXStages * kStageBytesA;
  float* s_bias = reinterpret_cast<float*>(w_base + (size_t)kWStages * kStageBytesB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint64_t* xfull = bars;                          // [kXStages]
  uint64_t* xempty = xfull + kXStages;             // [kXStages]
  uint64_t* wfull = xempty + kXStages;             // [kWStages]
  uint64_t* wempty = wfull + kWStages;             // [kWStages]
  uint64_t* tfull = wempty + kWStages;             // [2]
  uint64_t* tempty = tfull + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = p.kdim / kBK;
  const uint32_t seed = p.seed ^ (p.seed_offset ? __ldg(p.seed_offset) : 0u);

  if (threadIdx.x < 256) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == kEpiWarps && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    for (int i = 0; i < kXStages; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
    for (int i = 0; i < kWStages; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kEpiWarps) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      // x runs up to kXStages k-blocks ahead of the MMAs, W1 only kWStages: the two loads of a k-block are issued
      // when THEIR ring has room, x first (the long-latency one)
      const int total = ((p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * kblocks;
      int xi = 0, wi = 0;                                  // next k-block (flattened over this CTA's tiles) per ring
      while (wi < total || xi < total) {                    // either ring may be the one that still owes loads
        while (xi < total) {
          const int xs = xi % kXStages;
          if (!mbar_test(&xempty[xs], ((xi / kXStages) & 1) ^ 1)) break;
          const int tile = (int)blockIdx.x + (xi / kblocks) * (int)gridDim.x, kb = xi % kblocks;
          mbar_arrive_expect_tx(&xfull[xs], kStageBytesA);
          tma_load_2d(x_base + (size_t)xs * kStageBytesA, &tm_x, &xfull[xs], kb * kBK, tile * kBM);
          ++xi;
        }
        if (wi < total) {
          const int ws = wi % kWStages;
          if (mbar_test(&wempty[ws], ((wi / kWStages) & 1) ^ 1)) {
            mbar_arrive_expect_tx(&wfull[ws], kStageBytesB);
            tma_load_2d(w_base + (size_t)ws * kStageBytesB, &tm_w, &wfull[ws], (wi % kblocks) * kBK, 0);
            ++wi;
            continue;
          }
        }
        __nanosleep(20);
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, kD, 0, 0);
      int it = 0, n = 0;                                   // n: k-blocks issued so far (both rings advance together)
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kD;
        for (int kb = 0; kb < kblocks; ++kb, ++n) {
          const int xs = n % kXStages, ws = n % kWStages;
          mbar_wait(&xfull[xs], (n / kXStages) & 1);
          mbar_wait(&wfull[ws], (n / kWStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(x_base + (size_t)xs * kStageBytesA);
          const uint32_t sb = smem_u32(w_base + (size_t)ws * kStageBytesB);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            uint64_t ad = umma_desc_sw128(sa + k * 32, 0, 1024);
            uint64_t bd = umma_desc_sw128(sb + k * 32, 0, 1024);
            umma_f16(d_tmem, ad, bd, idesc, (kb | k) != 0);
          }
          umma_commit(&xempty[xs]);
          umma_commit(&wempty[ws]);
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ------------------------------ epilogue: TMEM -> bias/relu/dropout -> bf16 -> HBM ----
    const int quarter = warp & 3;         // TMEM lane quarter this warp may access
    const int half = warp >> 2;           // which kColsPerWarp of the 256 output columns
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const int row = tile * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.rows;
      bf16* out_row = p.h + (size_t)row * kD;
      // two 32-column chunks in flight: the TMEM load of the next chunk is issued before the current one is processed
      // (with one chunk at a time the ~300-cycle tcgen05.ld latency was exposed four times per tile and the epilogue,
      // not the tensor pipe, set the tile time)
      uint32_t vbuf[2][32];
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kD + half * kColsPerWarp;
      tmem_ld32(tbase, vbuf[0]);
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) {
        const int col0 = half * kColsPerWarp + cc * 32;
        tmem_ld_wait();
        if (cc + 1 < kChunks) tmem_ld32(tbase + (cc + 1) * 32, vbuf[(cc + 1) & 1]);
        const uint32_t (&v)[32] = vbuf[cc & 1];
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + col0 + j);
          float f[4] = {fmaxf(__uint_as_float(v[j]) + bb.x, 0.f), fmaxf(__uint_as_float(v[j + 1]) + bb.y, 0.f),
                        fmaxf(__uint_as_float(v[j + 2]) + bb.z, 0.f), fmaxf(__uint_as_float(v[j + 3]) + bb.w, 0.f)};
          if (p.drop_thresh) {
            // 16 hash bits per element (p is quantised to 1/65536): one 32-bit mix and one 32x32->64 multiply per four columns
            uint32_t z = (seed ^ ((uint32_t)row * 64u + (uint32_t)((col0 + j) >> 2))) * 0x9E3779B1u;
            z ^= z >> 15;
            z *= 0x85ebca6bu;
            z ^= z >> 13;
            const unsigned long long w = (unsigned long long)z * 0xD6E8FEB86659FD93ull;
            const uint32_t h0 = (uint32_t)w ^ (uint32_t)(w >> 32), h1 = (uint32_t)(w >> 29);
            const uint32_t bits[4] = {h0 & 0xffffu, h0 >> 16, h1 & 0xffffu, h1 >> 16};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              f[e] = (bits[e] >= p.drop_thresh) ? f[e] * p.keep_scale : 0.f;
          }
          packed[j / 2] = pack_bf16x2(f[0], f[1]);
          packed[j / 2 + 1] = pack_bf16x2(f[2], f[3]);
        }
        if (row_ok) {
          uint4* dst = reinterpret_cast<uint4*>(out_row + col0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------
// dW1 partials: out[ks][m][n] = sum_{r in K-split ks} dz[r][m] * x[r][n]
//   A = dz^T (MN-major: the 256 features are contiguous in memory), B = x^T (MN-major).
//   CTA (ks, mh, nh) owns a 128 x 256 fp32 accumulator in TMEM (256 columns).
// ------------------------------------------------------------------------------------------
constexpr int kDwStages = 4;
constexpr int kDwBK = 64;                                   // patch rows per stage
constexpr uint32_t kDwBytesA = kDwBK * 128 * 2;             // two [64 rows][64 feats] boxes = 16 KB
constexpr uint32_t kDwBytesB = kDwBK * 256 * 2;             // four boxes = 32 KB
constexpr uint32_t kDwStageBytes = kDwBytesA + kDwBytesB;
constexpr int kDwThreads = 6 * 32;
constexpr size_t kDwSmem = 1024 + (size_t)kDwStages * kDwStageBytes + 256;

struct DwParams {
  float* partial;     // (ksplit, 256, kin) fp32
  int rows;
  int kin;            // 512
  int ksplit;
  int kblocks_total;  // ceil(rows/64)
  uint32_t lbo, sbo;  // MN-major descriptor strides
};

__global__ void __launch_bounds__(kDwThreads, 1)
pathnet_dw_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_x,
                  const DwParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kDwStages * kDwStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kDwStages;
  uint64_t* tfull = bars + 2 * kDwStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile_n = p.kin / 256;
  const int ks = blockIdx.x / (2 * ntile_n);
  const int mh = (blockIdx.x / ntile_n) & 1;
  const int nh = blockIdx.x % ntile_n;
  const int per = (p.kblocks_total + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.kblocks_total, kb0 + per);
  const int nkb = max(0, kb1 - kb0);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_dz);
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], kDwStageBytes);
        uint8_t* sa = smem + (size_t)stage * kDwStageBytes;
#pragma unroll
        for (int b = 0; b < 2; ++b)
          tma_load_2d(sa + b * (kDwBK * 128), &tm_dz, &full[stage], mh * 128 + b * 64, kb * kDwBK);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          tma_load_2d(sa + kDwBytesA + b * (kDwBK * 128), &tm_x, &full[stage], nh * 256 + b * 64, kb * kDwBK);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 256, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)stage * kDwStageBytes);
        const uint32_t sb = sa + kDwBytesA;
#pragma unroll
        for (int k = 0; k < kDwBK / 16; ++k) {
          // MN-major: LBO = stride between 64-wide MN boxes (8 KB), SBO = 8 k-rows (1 KB)
          uint64_t ad = umma_desc_sw128(sa + k * 2048, p.lbo, p.sbo);
          uint64_t bd = umma_desc_sw128(sb + k * 2048, p.lbo, p.sbo);
          umma_f16(tmem_base, ad, bd, idesc, (i | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == kDwStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull);
    }
  } else {
    // epilogue: 4 warps, thread = one output feature row (m), 256 fp32 columns
    const int quarter = warp & 3;
    const int m = mh * 128 + quarter * 32 + lane;
    float* out = p.partial + ((size_t)ks * kD + m) * p.kin + nh * 256;
    if (nkb > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + cc * 32, v);
        tmem_ld_wait();
        uint4* dst = reinterpret_cast<uint4*>(out + cc * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) dst[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      }
    } else {
      for (int c = 0; c < 256; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0, 0, 0, 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// out[i] = sum_s partial[s][i]   (deterministic split-K / split-chunk reduction)
__global__ void sum_partials_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                    int nsplit, size_t n, float scale, int accumulate) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int s = 0; s < nsplit; ++s) {
    float4 v = *reinterpret_cast<const float4*>(partial + (size_t)s * n + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
  float4* o = reinterpret_cast<float4*>(out + i);
  if (accumulate) { float4 t = *o; acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w; }
  *o = acc;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host launchers (called from capi.cu)
// ------------------------------------------------------------------------------------------
int launch_pathnet_fwd(const bf16* x, const bf16* w1, const float* b1, bf16* h, int rows, int kdim,
                       float p_drop, uint32_t seed, cudaStream_t st) {
  if (rows <= 0) return IMP_OK;
  if (kdim % kBK != 0 || kdim <= 0) IMP_FAIL(IMP_ERR_ARG, "pathnet_fwd: kdim %d must be a positive multiple of 64", kdim);
  if (p_drop < 0.f || p_drop >= 1.f) IMP_FAIL(IMP_ERR_ARG, "pathnet_fwd: p_drop %f out of [0,1)", p_drop);
  CUtensorMap tm_x, tm_w;
  int rc;
  if ((rc = imp_make_tmap_2d(&tm_x, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kdim, rows, (uint64_t)kdim * 2, kBK, kBM,
                             CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = imp_make_tmap_2d(&tm_w, w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kdim, kD, (uint64_t)kdim * 2, kBK, kD,
                             CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  FwdParams p;
  p.bias = b1; p.h = h; p.rows = rows; p.kdim = kdim;
  p.num_tiles = (rows + kBM - 1) / kBM;
  p.drop_thresh = (uint32_t)(p_drop * 65536.f + 0.5f);
  p.keep_scale = p.drop_thresh ? 65536.f / (65536.f - (float)p.drop_thresh) : 1.f;
  p.seed = seed;
  p.seed_offset = imp_seed_offset_ptr();
  int grid = min(p.num_tiles, imp_num_sms());
  static const int cfg = []() { const char* e = getenv("IMP_PATHNET_RINGS"); return e ? atoi(e) : 0; }();   // tuning switch
#define IMP_PF(XS, WS)                                                                                               \
  do {                                                                                                               \
    constexpr size_t smem = fwd_smem<XS, WS>();                                                                      \
    static_assert(smem <= 227 * 1024, "pathnet_fwd shared memory");                                                  \
    { const int rc_ = imp_ensure_smem((const void*)pathnet_fwd_kernel<XS, WS>, smem); if (rc_) return rc_; }         \
    IMP_LAUNCH("pathnet_fwd", st, pathnet_fwd_kernel<XS, WS><<<grid, kFwdThreads, smem, st>>>(tm_x, tm_w, p));      \
  } while (0)
  switch (cfg) {
    case 1: IMP_PF(8, 2); break;
    case 2: IMP_PF(6, 3); break;
    case 3: IMP_PF(5, 3); break;
    case 4: IMP_PF(3, 4); break;
    default: IMP_PF(4, 4); break;
  }
#undef IMP_PF
  return IMP_OK;
}

int launch_sum_partials(const float* partial, float* out, int nsplit, size_t n, float scale, int accumulate,
                        cudaStream_t st) {
  if (n % 4) IMP_FAIL(IMP_ERR_ARG, "sum_partials: n must be a multiple of 4");
  size_t threads = n / 4;
  IMP_LAUNCH("sum_partials", st, sum_partials_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(partial, out, nsplit, n, scale, accumulate));
  return IMP_OK;
}

size_t pathnet_dw_workspace_bytes(int kin) {
  int ntile = 2 * (kin / 256);
  int ksplit = max(1, imp_num_sms() / ntile);
  return (size_t)ksplit * kD * kin * sizeof(float);
}

int launch_pathnet_dw(const bf16* dz, const bf16* x, float* dw, float* workspace, int rows, int kin,
                      int accumulate, cudaStream_t st) {
  if (kin % 256 != 0) IMP_FAIL(IMP_ERR_ARG, "pathnet_dw: in-features %d must be a multiple of 256", kin);
  if (rows <= 0) {
    if (!accumulate) IMP_CUDA(cudaMemsetAsync(dw, 0, (size_t)kD * kin * 4, st));
    return IMP_OK;
  }
  CUtensorMap tm_dz, tm_x;
  int rc;
  if ((rc = imp_make_tmap_2d(&tm_dz, dz, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kD, rows, (uint64_t)kD * 2, 64, kDwBK,
                             CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = imp_make_tmap_2d(&tm_x, x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, kin, rows, (uint64_t)kin * 2, 64, kDwBK,
                             CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  DwParams p;
  int ntile = 2 * (kin / 256);
  p.partial = workspace; p.rows = rows; p.kin = kin;
  p.kblocks_total = (rows + kDwBK - 1) / kDwBK;
  p.ksplit = max(1, min(imp_num_sms() / ntile, p.kblocks_total));
  p.lbo = kDwBK * 128; p.sbo = 1024;
  { const int rc_ = imp_ensure_smem((const void*)pathnet_dw_kernel, kDwSmem); if (rc_) return rc_; }
  IMP_LAUNCH("pathnet_dw", st, pathnet_dw_kernel<<<p.ksplit * ntile, kDwThreads, kDwSmem, st>>>(tm_dz, tm_x, p));
  return launch_sum_partials(workspace, dw, p.ksplit, (size_t)kD * kin, 1.f, accumulate, st);
}
