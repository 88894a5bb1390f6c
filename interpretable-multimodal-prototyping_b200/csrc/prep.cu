// A0  sentinel strip and the wire-format conversion of the reference batch layout.
// Reference: medmm/modeling/models/umeml_gan.py:401-410 (first row holding a -10000 element ends
// the bag) and medmm/data/data_manager.py:356-367,387 (bags padded to 10000 rows with -10000).
//
//   bag_lengths : img (B,Npad,D) fp32 -> len[b] = first row with ANY element == sentinel, else Npad
//   cu_seqlens  : exclusive prefix sum of len (B+1 entries), on the device (no host sync)
//   pack_bags   : valid rows -> packed (sum len, D) bf16, rows of bag b at [cu[b], cu[b+1])
#include "common.cuh"
#include "launchers.h"
#include <algorithm>

namespace {

__global__ void init_lengths_kernel(int* __restrict__ len, int B, int npad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) len[i] = npad;
}

// one warp per row: 128-bit coalesced loads, ballot on the sentinel compare, atomicMin per hit row
__global__ void bag_lengths_kernel(const float* __restrict__ img, int* __restrict__ len, long long total_rows,
                                   int npad, int d, float sentinel) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nvec = d >> 2;
  for (long long row = warp0; row < total_rows; row += nwarps) {
    const float4* src = reinterpret_cast<const float4*>(img + row * d);
    bool hit = false;
    for (int c = lane; c < nvec; c += 32) {
      const float4 v = __ldg(src + c);
      hit |= (v.x == sentinel) | (v.y == sentinel) | (v.z == sentinel) | (v.w == sentinel);
    }
    for (int c = (nvec << 2) + lane; c < d; c += 32) hit |= img[row * d + c] == sentinel;
    if (__any_sync(0xffffffffu, hit) && lane == 0) atomicMin(len + (int)(row / npad), (int)(row % npad));
  }
}

// single block: cu[0] = 0, cu[b+1] = cu[b] + len[b]
__global__ void cu_seqlens_kernel(const int* __restrict__ len, int* __restrict__ cu, int B) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int per = (B + nt - 1) / nt;
  const int b0 = tid * per, b1 = min(B, b0 + per);
  int sum = 0;
  for (int b = b0; b < b1; ++b) sum += len[b];
  s_part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < nt; ++i) { int v = s_part[i]; s_part[i] = run; run += v; }
    cu[B] = run;
  }
  __syncthreads();
  int run = s_part[tid];
  for (int b = b0; b < b1; ++b) { cu[b] = run; run += len[b]; }
}

// grid (row chunks, B): fp32 valid rows -> bf16 packed rows
__global__ void pack_bags_kernel(const float* __restrict__ img, const int* __restrict__ cu, bf16* __restrict__ out,
                                 int npad, int d) {
  const int b = blockIdx.y;
  const int begin = cu[b], n = cu[b + 1] - begin;
  const int nvec = d >> 2;                                    // d % 4 == 0 (checked on the host)
  const long long items = (long long)n * nvec;
  const float4* src = reinterpret_cast<const float4*>(img + (size_t)b * npad * d);
  uint2* dst = reinterpret_cast<uint2*>(out + (size_t)begin * d);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    dst[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

__global__ void cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    reinterpret_cast<uint2*>(dst)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

}  // namespace

int launch_bag_lengths(const float* img, int B, int npad, int d, float sentinel, int* lengths, int* cu, cudaStream_t st) {
  if (B <= 0 || npad <= 0 || d <= 0) IMP_FAIL(IMP_ERR_ARG, "bag_lengths: bad shape (%d,%d,%d)", B, npad, d);
  if (d % 4 != 0) IMP_FAIL(IMP_ERR_ARG, "bag_lengths: feature dim %d must be a multiple of 4", d);
  if ((long long)B * npad >= (1LL << 31)) IMP_FAIL(IMP_ERR_ARG, "bag_lengths: %lld rows exceed int32 offsets", (long long)B * npad);
  IMP_LAUNCH("init_lengths", st, init_lengths_kernel<<<(B + 255) / 256, 256, 0, st>>>(lengths, B, npad));
  const long long rows = (long long)B * npad;
  const int blocks = (int)std::min<long long>((rows + 7) / 8, (long long)imp_num_sms() * 16);
  IMP_LAUNCH("bag_lengths", st, bag_lengths_kernel<<<blocks, 256, 0, st>>>(img, lengths, rows, npad, d, sentinel));
  if (cu) {
    IMP_LAUNCH("cu_seqlens", st, cu_seqlens_kernel<<<1, 1024, 0, st>>>(lengths, cu, B));
  }
  return IMP_OK;
}

int launch_pack_bags(const float* img, int B, int npad, int d, const int* cu, bf16* out, cudaStream_t st) {
  if (B <= 0 || npad <= 0 || d <= 0 || d % 4 != 0) IMP_FAIL(IMP_ERR_ARG, "pack_bags: bad shape (%d,%d,%d)", B, npad, d);
  const int chunks = max(1, min(64, (2 * imp_num_sms() * 4 + B - 1) / B));
  IMP_LAUNCH("pack_bags", st, pack_bags_kernel<<<dim3(chunks, B), 256, 0, st>>>(img, cu, out, npad, d));
  return IMP_OK;
}

int launch_cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st) {
  if (n % 4 != 0) IMP_FAIL(IMP_ERR_ARG, "cast_bf16: element count must be a multiple of 4");
  if (n == 0) return IMP_OK;
  const size_t n4 = n / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, (size_t)imp_num_sms() * 16);
  IMP_LAUNCH("cast_bf16", st, cast_bf16_kernel<<<blocks, 256, 0, st>>>(src, dst, n4));
  return IMP_OK;
}
