// microbenchmark (nvcc -gencode arch=compute_100a,code=sm_100a -O3; results in profiles/r01_microbench.md): shared-memory integer atomic add (no return) vs LDS+FADD+STS, per warp instruction
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int* out, int iters, long long* cyc) {
  extern __shared__ int s[];
  for (int i = threadIdx.x; i < 40 * 128 * 4; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int row = threadIdx.x & 127;           // MODE 0/1: 4 warps share a row (atomics) ; MODE 2: private column
  int idx = (threadIdx.x * 7) & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      idx = (idx * 5 + 3 + u) & 31;
      if (MODE == 0) atomicAdd(&s[idx * 128 + row], it + u);
      else if (MODE == 1) atomicAdd(&s[idx * 512 + threadIdx.x], it + u);
      else { int* p = &s[idx * 512 + threadIdx.x]; *p = *p + it + u; }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int acc = 0;
  for (int i = threadIdx.x; i < 40 * 128 * 4; i += blockDim.x) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  int* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920); cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920); cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 81920);
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 512, 81920>>>(out, iters, cyc);
      if (mode == 1) k<1><<<148, 512, 81920>>>(out, iters, cyc);
      if (mode == 2) k<2><<<148, 512, 81920>>>(out, iters, cyc);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double per = (double)h[0] / (iters * 8.0 * 16);   // cycles per warp-instruction slot (16 warps per SM)
    printf("mode %d: %.2f SM cycles per warp update (16 warps), err %s\n", mode, per, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
