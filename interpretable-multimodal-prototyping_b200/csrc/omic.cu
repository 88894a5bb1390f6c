// A7  per-pathway genomic encoders and A8 masked handling of missing omics.
// Reference: medmm/modeling/models/umeml_gan.py:274-283,413-419 (six Linear+ReLU+Dropout encoders on
// gathered gene groups) and :380-392,500-511 (mean imputation of masked genes; sample-level
// replacement and batch-ratio blend of the omic tokens with the generator output).
//
// The encoders are skinny GEMMs (M = batch, K = G_k <= 1538, N = 256): bound by reading the
// 3.4 MB of weights once, so they run as warp-per-output-feature dot products with the gather
// and the imputation fused into the operand staging -- no tensor-core tile would be filled.
#include "common.cuh"
#include "launchers.h"
#include <algorithm>

namespace {

constexpr int kD = 256;
constexpr int kMaxGroups = 8;
constexpr int kBB = 8;            // batch rows per CTA

struct OmicParams {
  const float* x;                 // (B, G)
  const int* mask;                // (B, G) int32 or null: 1 = gene missing -> use means
  const float* means;             // (G) or null
  const int* idx;                 // concatenated gather indices
  const float* w[kMaxGroups];     // (256, G_k)
  const float* bias[kMaxGroups];  // (256)
  int goff[kMaxGroups + 1];
  float* out;                     // (B, K, 256)
  int B, G, K;
  float keep_scale;
  uint32_t drop_thresh, seed;
  const uint32_t* seed_offset;    // device word XOR-ed into the seed, or null
};

__device__ __forceinline__ float gather_gene(const OmicParams& p, int b, int gene) {
  const size_t o = (size_t)b * p.G + gene;
  if (p.mask && p.means && p.mask[o] != 0) return __ldg(p.means + gene);     // umeml_gan.py:391-392
  return __ldg(p.x + o);
}

__global__ void __launch_bounds__(256) omic_fwd_kernel(const OmicParams p) {
  extern __shared__ float xs[];                         // [kBB][Gk]
  const int k = blockIdx.y, b0 = blockIdx.z * kBB;
  const int g0 = p.goff[k], gk = p.goff[k + 1] - g0;
  const int nb = min(kBB, p.B - b0);
  for (int i = threadIdx.x; i < kBB * gk; i += blockDim.x) {
    const int bb = i / gk, g = i % gk;
    xs[i] = bb < nb ? gather_gene(p, b0 + bb, __ldg(p.idx + g0 + g)) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * 8 + warp;
  const float* wrow = p.w[k] + (size_t)f * gk;
  float acc[kBB];
#pragma unroll
  for (int bb = 0; bb < kBB; ++bb) acc[bb] = 0.f;
  for (int g = lane; g < gk; g += 32) {
    const float wv = __ldg(wrow + g);
#pragma unroll
    for (int bb = 0; bb < kBB; ++bb) acc[bb] += wv * xs[bb * gk + g];
  }
#pragma unroll
  for (int bb = 0; bb < kBB; ++bb) acc[bb] = warp_sum(acc[bb]);
  if (lane < nb) {
    float v = 0.f;
#pragma unroll
    for (int bb = 0; bb < kBB; ++bb) if (lane == bb) v = acc[bb];
    v = fmaxf(v + __ldg(p.bias[k] + f), 0.f);
    const int b = b0 + lane;
    if (p.drop_thresh) {
      const uint32_t seed = p.seed ^ (p.seed_offset ? __ldg(p.seed_offset) : 0u);
      const uint32_t hsh = mix32(seed ^ (((uint32_t)(b * p.K + k) * 256u + (uint32_t)f) * 0x9E3779B1u));
      v = ((hsh & 0xffffu) >= p.drop_thresh) ? v * p.keep_scale : 0.f;
    }
    p.out[((size_t)b * p.K + k) * kD + f] = v;
  }
}

struct OmicBwdParams {
  OmicParams f;                   // inputs of the forward (x, mask, means, idx, goff, B, G, K, keep_scale)
  const float* out;               // (B,K,256) forward output (relu/dropout mask)
  const float* dout;              // (B,K,256)
  float* dw[kMaxGroups];          // (256, G_k)
  float* db[kMaxGroups];          // (256)
  int accumulate;
};

// CTA = (32 genes of group k) x all 256 features; thread (fq, lane): gene g0+lane, features fq*32..+31
__global__ void __launch_bounds__(256) omic_bwd_kernel(const OmicBwdParams p) {
  __shared__ float s_dz[kBB][kD];
  __shared__ float s_x[kBB][32];
  const int k = blockIdx.y;
  const int goff = p.f.goff[k], gk = p.f.goff[k + 1] - goff;
  const int gbase = blockIdx.x * 32;
  if (gbase >= gk) return;
  const int fq = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = gbase + lane;
  const int gene = g < gk ? __ldg(p.f.idx + goff + g) : 0;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  float dbacc = 0.f;
  for (int b0 = 0; b0 < p.f.B; b0 += kBB) {
    const int nb = min(kBB, p.f.B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < kBB * kD; i += 256) {
      const int bb = i / kD, f = i % kD;
      float v = 0.f;
      if (bb < nb) {
        const size_t o = ((size_t)(b0 + bb) * p.f.K + k) * kD + f;
        v = p.out[o] > 0.f ? p.dout[o] * p.f.keep_scale : 0.f;
      }
      s_dz[bb][f] = v;
    }
    if (threadIdx.x < kBB * 32) {
      const int bb = threadIdx.x >> 5, l = threadIdx.x & 31;
      const int gg = gbase + l;
      s_x[bb][l] = (bb < nb && gg < gk) ? gather_gene(p.f, b0 + bb, __ldg(p.f.idx + goff + gg)) : 0.f;
    }
    __syncthreads();
    for (int bb = 0; bb < nb; ++bb) {
      const float xv = s_x[bb][lane];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] += s_dz[bb][fq * 32 + i] * xv;
      if (blockIdx.x == 0) dbacc += s_dz[bb][threadIdx.x];
    }
  }
  (void)gene;
  if (g < gk) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float* dst = p.dw[k] + (size_t)(fq * 32 + i) * gk + g;
      *dst = p.accumulate ? *dst + acc[i] : acc[i];
    }
  }
  if (blockIdx.x == 0) {
    float* dst = p.db[k] + threadIdx.x;
    *dst = p.accumulate ? *dst + dbacc : dbacc;
  }
}

// A8 (ii)/(iii): h_omic <- where(without_omic[b], gen, h_omic); then (1-r) h_omic + r gen with
// r = sum(insample mask)/numel over the WHOLE batch (umeml_gan.py:503-511).  r is reduced on the
// device: the reference's "if sum > 0" host syncs become r == 0 / no-sample no-ops.
__global__ void mask_ratio_kernel(const int* __restrict__ mask, long long n, float* __restrict__ ratio_sum) {
  long long local = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    local += mask[i];
  float v = warp_sum((float)local);
  if ((threadIdx.x & 31) == 0 && v != 0.f) atomicAdd(ratio_sum, v);
}
__global__ void omic_blend_kernel(const float* __restrict__ h, const float* __restrict__ gen,
                                  const int* __restrict__ without, const float* __restrict__ ratio_sum,
                                  float inv_numel, float* __restrict__ out, float* __restrict__ ratio_out,
                                  int B, int per_sample) {
  const float r = ratio_sum ? *ratio_sum * inv_numel : 0.f;
  if (ratio_out && blockIdx.x == 0 && threadIdx.x == 0) *ratio_out = r;
  const long long n = (long long)B * per_sample;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const float gv = gen[i];
    float hv = (without && without[b] == 1) ? gv : h[i];
    out[i] = (1.f - r) * hv + r * gv;
  }
}

}  // namespace

static int fill_params(OmicParams& p, const float* x, const int* mask, const float* means, const int* idx,
                       const int* group_offsets, int K, int B, int G, float p_drop, unsigned seed) {
  if (K < 1 || K > kMaxGroups) IMP_FAIL(IMP_ERR_ARG, "omic: %d gene groups (1..%d supported)", K, kMaxGroups);
  if (B <= 0 || G <= 0) IMP_FAIL(IMP_ERR_ARG, "omic: bad shape B=%d G=%d", B, G);
  if (p_drop < 0.f || p_drop >= 1.f) IMP_FAIL(IMP_ERR_ARG, "omic: p_drop %f out of [0,1)", p_drop);
  p.x = x; p.mask = mask; p.means = means; p.idx = idx; p.B = B; p.G = G; p.K = K; p.seed = seed; p.seed_offset = imp_seed_offset_ptr();
  for (int k = 0; k <= K; ++k) p.goff[k] = group_offsets[k];
  for (int k = 0; k < K; ++k)
    if (p.goff[k + 1] <= p.goff[k]) IMP_FAIL(IMP_ERR_ARG, "omic: group %d is empty", k);
  p.drop_thresh = (uint32_t)(p_drop * 65536.f + 0.5f);
  p.keep_scale = p.drop_thresh ? 65536.f / (65536.f - (float)p.drop_thresh) : 1.f;
  return IMP_OK;
}

int launch_omic_fwd(const float* x, const int* mask, const float* means, const int* idx, const int* group_offsets,
                    int K, const float* const* w, const float* const* bias, int B, int G, float p_drop, unsigned seed,
                    float* out, cudaStream_t st) {
  OmicParams p;
  int rc = fill_params(p, x, mask, means, idx, group_offsets, K, B, G, p_drop, seed);
  if (rc) return rc;
  int gmax = 0;
  for (int k = 0; k < K; ++k) { p.w[k] = w[k]; p.bias[k] = bias[k]; gmax = std::max(gmax, p.goff[k + 1] - p.goff[k]); }
  p.out = out;
  const size_t smem = (size_t)kBB * gmax * sizeof(float);
  if (smem > 200 * 1024) IMP_FAIL(IMP_ERR_ARG, "omic: gene group of %d genes exceeds the staging buffer", gmax);
  { const int rc_ = imp_ensure_smem((const void*)omic_fwd_kernel, smem); if (rc_) return rc_; }
  IMP_LAUNCH("omic_fwd", st, omic_fwd_kernel<<<dim3(kD / 8, K, (B + kBB - 1) / kBB), 256, smem, st>>>(p));
  return IMP_OK;
}

int launch_omic_bwd(const float* x, const int* mask, const float* means, const int* idx, const int* group_offsets,
                    int K, int B, int G, float p_drop, const float* out, const float* dout, float* const* dw,
                    float* const* db, int accumulate, cudaStream_t st) {
  OmicBwdParams p;
  int rc = fill_params(p.f, x, mask, means, idx, group_offsets, K, B, G, p_drop, 0);
  if (rc) return rc;
  int gmax = 0;
  for (int k = 0; k < K; ++k) { p.dw[k] = dw[k]; p.db[k] = db[k]; gmax = std::max(gmax, p.f.goff[k + 1] - p.f.goff[k]); }
  p.out = out; p.dout = dout; p.accumulate = accumulate;
  IMP_LAUNCH("omic_bwd", st, omic_bwd_kernel<<<dim3((gmax + 31) / 32, K), 256, 0, st>>>(p));
  return IMP_OK;
}

int launch_omic_blend(const float* h_omic, const float* h_gen, const int* without_omic, const int* insample_mask,
                      long long mask_numel, int B, int per_sample, float* scratch, float* out, float* ratio_out,
                      cudaStream_t st) {
  if (B <= 0 || per_sample <= 0) IMP_FAIL(IMP_ERR_ARG, "omic_blend: bad shape");
  float* ratio_sum = nullptr;
  if (insample_mask && mask_numel > 0) {
    if (!scratch) IMP_FAIL(IMP_ERR_ARG, "omic_blend: scratch (1 float) required with an in-sample mask");
    ratio_sum = scratch;
    IMP_CUDA(cudaMemsetAsync(ratio_sum, 0, sizeof(float), st));
    const int blocks = (int)std::min<long long>((mask_numel + 255) / 256, (long long)imp_num_sms() * 4);
    IMP_LAUNCH("mask_ratio", st, mask_ratio_kernel<<<blocks, 256, 0, st>>>(insample_mask, mask_numel, ratio_sum));
  }
  const long long n = (long long)B * per_sample;
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)imp_num_sms() * 4);
  IMP_LAUNCH("omic_blend", st, omic_blend_kernel<<<blocks, 256, 0, st>>>(h_omic, h_gen, without_omic, ratio_sum,
                                            mask_numel > 0 ? 1.f / (float)mask_numel : 0.f, out, ratio_out, B, per_sample));
  return IMP_OK;
}
