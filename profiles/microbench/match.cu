// microbenchmark (nvcc -gencode arch=compute_100a,code=sm_100a -O3; results in profiles/r01_microbench.md): warp-level segmented sum by key
// (match.any + redux.sync.add + leader atomic), the column-side accumulation a symmetric pair sweep would need
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(int* out, int iters, long long* cyc, int distinct) {
  extern __shared__ int s[];
  for (int i = threadIdx.x; i < 64 * 41; i += blockDim.x) s[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int key = (lane * 7 + warp) % distinct;
  int acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      key = (key * 5 + 3 + u) % distinct;
      const int val = it + u + lane;
      if (MODE == 0) {                       // match only
        acc += __match_any_sync(0xffffffffu, key);
      } else if (MODE == 1) {                // match + segmented redux + leader atomic
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int sum = __reduce_add_sync(m, val);
        if (lane == __ffs(m) - 1) atomicAdd(&s[((warp * 4 + u) & 63) * 41 + key], sum);
      } else {                               // plain per-lane atomic (collisions serialise)
        atomicAdd(&s[((warp * 4 + u) & 63) * 41 + key], val);
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  for (int i = threadIdx.x; i < 64 * 41; i += blockDim.x) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
  int* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  const char* names[] = {"match.any", "match+redux+leader ATOMS", "per-lane ATOMS"};
  for (int mode = 0; mode < 3; ++mode) for (int d : {32, 8, 2}) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 512, 64 * 41 * 4>>>(out, iters, cyc, d);
      if (mode == 1) k<1><<<148, 512, 64 * 41 * 4>>>(out, iters, cyc, d);
      if (mode == 2) k<2><<<148, 512, 64 * 41 * 4>>>(out, iters, cyc, d);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-26s distinct keys %2d: %.2f SMSP cycles per warp-level update (4 warps per SMSP) %s\n", names[mode], d,
           (double)h[0] / (iters * 8.0 * 4), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
