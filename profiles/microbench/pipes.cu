// microbenchmark (nvcc -gencode arch=compute_100a,code=sm_100a -O3; results in profiles/r01_microbench.md): issue rate per SM sub-partition of the sweep's main instructions (4 warps / SMSP, 8 chains)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r; asm volatile("{.reg .b64 a,b,c; mov.b64 a,{%2,%3}; mov.b64 b,{%4,%5}; add.rn.f32x2 c,a,b; mov.b64 {%0,%1},c;}" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y)); return r; }
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc, float seed) {
  float2 a[8]; float m[8];
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed * i + threadIdx.x, seed + i); m[i] = seed * (i + 1); }
  const float2 inc = make_float2(seed, seed * 0.5f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE == 2) a[i] = add2(a[i], inc);
      if (MODE == 1 || MODE == 2) m[i] = fminf(fminf(a[i].x, a[(i + 1) & 7].y), m[i]);
      if (MODE == 3) m[i] = fminf(a[i].x, m[i]) + 0.f * a[i].y;     // FMNMX 2-input + a dependent touch
      if (MODE == 4) { int v = __float_as_int(m[i]); v = __viaddmin_s32(v, __float_as_int(a[i].x), __float_as_int(a[i].y)); m[i] = __int_as_float(v); }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y + m[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  const char* names[] = {"FADD2 only", "FMNMX3 only", "FADD2+FMNMX3", "FMNMX+FFMA", "VIADDMNMX"};
  for (int mode = 0; mode < 5; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<148, 512>>>(out, iters, cyc, 1.0001f);
      if (mode == 1) k<1><<<148, 512>>>(out, iters, cyc, 1.0001f);
      if (mode == 2) k<2><<<148, 512>>>(out, iters, cyc, 1.0001f);
      if (mode == 3) k<3><<<148, 512>>>(out, iters, cyc, 1.0001f);
      if (mode == 4) k<4><<<148, 512>>>(out, iters, cyc, 1.0001f);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-14s: %.2f cycles per loop body of 8 (4 warps per SMSP -> %.2f SMSP cycles per warp body-instruction group)\n", names[mode], (double)h[0] / iters, (double)h[0] / iters / 8.0 / 4.0 * 1.0);
  }
  return 0;
}
