// microbenchmark (nvcc -gencode arch=compute_100a,code=sm_100a -O3; results in profiles/r01_microbench.md): shared int atomics with k distinct addresses per warp instruction (same-address collisions)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* out, int iters, long long* cyc, int distinct, int stride) {
  extern __shared__ int s[];
  for (int i = threadIdx.x; i < 64 * 41; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int key = lane % distinct;                   // token index of this lane
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int col = (warp * 4 + u + it) & 63;          // same column for the whole warp
      atomicAdd(&s[col * stride + ((key + u) % 40)], it + u);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  int acc = 0;
  for (int i = threadIdx.x; i < 64 * 41; i += blockDim.x) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  int* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  int ds[] = {32, 16, 8, 4, 2, 1};
  for (int stride : {41, 40}) for (int d : ds) {
    for (int rep = 0; rep < 2; ++rep) { k<<<148, 512, 64 * 41 * 4>>>(out, iters, cyc, d, stride); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("stride %d distinct %2d: %.2f SM cycles per warp atomic (16 warps) %s\n", stride, d, (double)h[0] / (iters * 8.0 * 16), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
