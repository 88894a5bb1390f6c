// N1  Nystrom attention core of the token-level tail (reference: medmm/modeling/ops/attention.py:105-127,
//     moore_penrose_iter_pinv ops/utils.py:116-131), on the reduced (n+1) x (n+1) matrices of token_tail.nystrom_short.
//
//   forward :  Z_0 = s A^T ;  Z_{k+1} = 1/4 Z_k (13 I - A Z_k (15 I - A Z_k (7 I - A Z_k)))   (k < iters)
//              y   = rows 1..n of  A (Z_iters (A [0 ; v]))
//   backward:  dA, ds (one partial per matrix), dv  from dy, with the Z_k recomputed in shared memory
//
// A is (n+1) x (n+1) with n <= 47 tokens, v is n x d (d = 32 or 64): ~30 matrix products of 41^3 per (slide, head)
// forward and ~100 backward.  As batched library calls that is ~85 GEMM launches and ~100 element-wise launches per
// attention layer and training step, seven layers per step: most of the token tail's launches.  Here one CTA owns
// one (slide, head), keeps every matrix in shared memory and runs the whole chain: two launches per layer and step.
// fp32 FMA on 2 x 2 register tiles (the products are far too small for the tensor core and the iteration wants fp32).
#include "common.cuh"
#include "launchers.h"

namespace {

constexpr int kNyThreads = 512;        // four groups of 128 threads
constexpr int kNyMaxN = 48;          // n + 1 rounded up to a multiple of 4

// C (n x m) = diag * I_{ndiag} + alpha * op(A) op(B) (+ C when ACC); op(A) is n x kd, op(B) is kd x m.
// n and m are multiples of 4; every matrix is zero outside its real rows / columns, so no bounds are checked.
// 4 x 4 register tiles: 8 shared-memory loads per 16 FMAs (2 x 2 tiles, one load per FMA: backward 224 -> 197 us, forward
// 69 -> 51 us per layer at 32 slides x 8 heads, 39 tokens).  A software-pipelined variant (register ping-pong, 256
// threads, 255 registers, one CTA per SM) measured slower (232 / 92 us): the products are bound by shared-memory
// wavefronts (12 per k step and warp with the 44-word rows wrapping the 32 banks), not by load latency.
template <bool TA, bool TB, bool ACC>
__device__ __forceinline__ void mm(float* __restrict__ C, int ldc, const float* __restrict__ A, int lda,
                                   const float* __restrict__ B, int ldb, int n, int m, int kd, float alpha,
                                   float diag = 0.f, int ndiag = 0, int grp = 0) {
  // one group of 128 threads per product (<= 144 tiles): independent products of a phase run side by side on
  // different groups and hide each other's shared-memory latency
  if ((int)(threadIdx.x >> 7) != grp) return;
  const int tm = m >> 2, tiles = (n >> 2) * tm;
  for (int t = threadIdx.x & 127; t < tiles; t += 128) {
    const int i = (t / tm) * 4, j = (t % tm) * 4;
    float c[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[r][q] = 0.f;
    const float* ap = TA ? A + i : A + i * lda;
    const float* bp = TB ? B + j * ldb : B + j;
#pragma unroll 2
    for (int k = 0; k < kd; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = TA ? ap[k * lda + r] : ap[r * lda + k];
#pragma unroll
      for (int q = 0; q < 4; ++q) b[q] = TB ? bp[q * ldb + k] : bp[k * ldb + q];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[r][q] = fmaf(a[r], b[q], c[r][q]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float* crow = C + (i + r) * ldc + j;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v = alpha * c[r][q];
        if (i == j && r == q && i + r < ndiag) v += diag;      // 4 x 4 tiles are aligned with the diagonal
        if (ACC) v += crow[q];
        crow[q] = v;
      }
    }
  }
}

struct NyParams {
  const float* mat;        // (BH, N, N)
  const float* inv_scale;  // (1) device scalar s
  const float* v;          // (BH, n, d)
  const float* dy;         // (BH, n, d)            backward only
  float* y;                // (BH, n, d)            forward only
  float* dmat;             // (BH, N, N)
  float* dv;               // (BH, n, d)
  float* dscale;           // (BH) partial derivatives wrt s
  float* zs;               // (BH, iters + 1, NP * (NP + 1)) the iterates Z_k in the shared-memory layout: written by the
                           // forward (may be null), read by the backward instead of running the iteration again
  const float* conv_w;     // (heads, taps) depth-wise residual convolution over the tokens (attention.py:129-131), or null
  float* dconv;            // (BH, taps) per-matrix partial gradients of conv_w        backward only
  int heads, taps;         // matrix index = slide * heads + head; taps odd (33)
  int N, d, iters;
};

// A <- mat (zero padded), Z0 <- s A^T
__device__ __forceinline__ void load_mat(const NyParams& p, float* A, float* Z0, int NP, int LD, float s) {
  const float* src = p.mat + (size_t)blockIdx.x * p.N * p.N;
  for (int idx = threadIdx.x; idx < NP * NP; idx += kNyThreads) {
    const int i = idx / NP, j = idx - i * NP;
    const float a = (i < p.N && j < p.N) ? __ldg(src + i * p.N + j) : 0.f;
    A[i * LD + j] = a;
    Z0[j * LD + i] = s * a;
  }
}
// V1 <- [0 ; v] (row 0 and the padding rows are zero)
__device__ __forceinline__ void load_rows(const float* src, float* V, int N, int NP, int d, int LDV) {
  for (int idx = threadIdx.x; idx < NP * d; idx += kNyThreads) {
    const int i = idx / d, c = idx - i * d;
    V[i * LDV + c] = (i >= 1 && i < N) ? __ldg(src + (size_t)(i - 1) * d + c) : 0.f;
  }
}
// one step of the iteration: Znext = 1/4 Z (13 I - AZ (15 I - AZ (7 I - AZ))); AZ, T1, T2 are scratch and hold
// A Z, 7 I - A Z (overwritten by T3 = 13 I - ...) and 15 I - AZ T1 afterwards when KEEP (backward) is set
__device__ __forceinline__ void ns_terms(const float* A, const float* Z, float* AZ, float* T1, float* T2, float* T3,
                                         int N, int NP, int LD) {
  mm<false, false, false>(AZ, LD, A, LD, Z, LD, NP, NP, NP, 1.f);
  __syncthreads();
  for (int idx = threadIdx.x; idx < NP * NP; idx += kNyThreads) {
    const int i = idx / NP, j = idx - i * NP;
    T1[i * LD + j] = ((i == j && i < N) ? 7.f : 0.f) - AZ[i * LD + j];
  }
  __syncthreads();
  mm<false, false, false>(T2, LD, AZ, LD, T1, LD, NP, NP, NP, -1.f, 15.f, N);
  __syncthreads();
  mm<false, false, false>(T3, LD, AZ, LD, T2, LD, NP, NP, NP, -1.f, 13.f, N);
  __syncthreads();
}

__global__ void __launch_bounds__(kNyThreads, 2) nystrom_core_fwd_kernel(const NyParams p) {
  extern __shared__ float ny_smem[];
  const int N = p.N, NP = (N + 3) & ~3, LD = NP + 1, d = p.d, LDV = d + 1;
  const int MS = NP * LD;
  float* A = ny_smem;
  float* Z = A + MS;
  float* Zn = Z + MS;
  float* AZ = Zn + MS;
  float* T1 = AZ + MS;
  float* T2 = T1 + MS;
  float* T3 = T2 + MS;
  float* V1 = T3 + MS;
  float* W1 = V1 + NP * LDV;
  float* W2 = W1 + NP * LDV;
  const float s = __ldg(p.inv_scale);
  load_mat(p, A, Z, NP, LD, s);
  load_rows(p.v + (size_t)blockIdx.x * (N - 1) * d, V1, N, NP, d, LDV);
  __syncthreads();
  float* zdst = p.zs ? p.zs + (size_t)blockIdx.x * (p.iters + 1) * MS : nullptr;
  auto keep = [&](const float* Zk, int k) {          // the padding column is never read: copy the buffer as it is
    if (zdst) for (int idx = threadIdx.x; idx < MS; idx += kNyThreads) zdst[(size_t)k * MS + idx] = Zk[idx];
  };
  keep(Z, 0);
  for (int k = 0; k < p.iters; ++k) {
    ns_terms(A, Z, AZ, T1, T2, T3, N, NP, LD);
    mm<false, false, false>(Zn, LD, Z, LD, T3, LD, NP, NP, NP, 0.25f);
    __syncthreads();
    float* t = Z; Z = Zn; Zn = t;
    keep(Z, k + 1);
  }
  mm<false, false, false>(W1, LDV, A, LD, V1, LDV, NP, d, NP, 1.f);
  __syncthreads();
  mm<false, false, false>(W2, LDV, Z, LD, W1, LDV, NP, d, NP, 1.f);
  __syncthreads();
  mm<false, false, false>(W1, LDV, A, LD, W2, LDV, NP, d, NP, 1.f);        // W1 is free again: Y
  __syncthreads();
  float* dst = p.y + (size_t)blockIdx.x * (N - 1) * d;
  const float* cw = p.conv_w ? p.conv_w + (size_t)(blockIdx.x % p.heads) * p.taps : nullptr;
  const int half = p.taps >> 1;
  for (int idx = threadIdx.x; idx < (N - 1) * d; idx += kNyThreads) {
    const int i = idx / d, c = idx - i * d;
    float acc = W1[(i + 1) * LDV + c];
    if (cw) {                                       // + sum_t w[t] v[i + t - half]: the zero tokens the reference pads in
      const int t0 = max(0, half - i), t1 = min(p.taps, N - 1 - i + half);      // front act like the conv's own padding
      for (int t = t0; t < t1; ++t) acc = fmaf(__ldg(cw + t), V1[(i + t - half + 1) * LDV + c], acc);
    }
    dst[idx] = acc;
  }
}

__global__ void __launch_bounds__(kNyThreads, 1) nystrom_core_bwd_kernel(const NyParams p) {
  extern __shared__ float ny_smem[];
  const int N = p.N, NP = (N + 3) & ~3, LD = NP + 1, d = p.d, LDV = d + 1;
  const int MS = NP * LD, VS = NP * LDV;
  float* A = ny_smem;
  float* Zs = A + MS;                            // Z_0 .. Z_iters
  float* G = Zs + (size_t)(p.iters + 1) * MS;    // dZ_{k+1}
  float* Gn = G + MS;                            // dZ_k
  float* dA = Gn + MS;
  float* AZ = dA + MS;                           // scratch of the iteration; the final stage's vectors alias it
  float* T1 = AZ + MS;
  float* T2 = T1 + MS;
  float* T3 = T2 + MS;
  float* dAZ = T3 + MS;
  float* X1 = dAZ + MS;
  float* X2 = X1 + MS;
  // the six NP x d vectors of the final stage live inside the seven scratch matrices when they fit (the usual case:
  // 38 + 1 tokens, d = 32), behind them otherwise (a handful of tokens)
  float* V1 = (6 * VS <= 7 * MS) ? AZ : X2 + MS;
  float* W1 = V1 + VS;
  float* W2 = W1 + VS;
  float* dY = W2 + VS;
  float* dW1 = dY + VS;
  float* dW2 = dW1 + VS;
  const float s = __ldg(p.inv_scale);
  load_mat(p, A, Zs, NP, LD, s);
  __syncthreads();
  // ---- the iterates Z_k: from the forward launch when it kept them, else by running the iteration again ----
  if (p.zs) {
    const float* zsrc = p.zs + (size_t)blockIdx.x * (p.iters + 1) * MS;
    for (int idx = threadIdx.x; idx < (p.iters + 1) * MS; idx += kNyThreads) Zs[idx] = __ldg(zsrc + idx);
    __syncthreads();
  } else
  for (int k = 0; k < p.iters; ++k) {
    const float* Z = Zs + (size_t)k * MS;
    ns_terms(A, Z, AZ, T1, T2, T3, N, NP, LD);
    mm<false, false, false>(Zs + (size_t)(k + 1) * MS, LD, Z, LD, T3, LD, NP, NP, NP, 0.25f);
    __syncthreads();
  }
  const float* Zl = Zs + (size_t)p.iters * MS;
  // ---- y = A (Z (A V1)) ----
  load_rows(p.v + (size_t)blockIdx.x * (N - 1) * d, V1, N, NP, d, LDV);
  load_rows(p.dy + (size_t)blockIdx.x * (N - 1) * d, dY, N, NP, d, LDV);
  __syncthreads();
  mm<false, false, false>(W1, LDV, A, LD, V1, LDV, NP, d, NP, 1.f);
  __syncthreads();
  mm<false, false, false>(W2, LDV, Zl, LD, W1, LDV, NP, d, NP, 1.f);
  mm<true, false, false>(dW2, LDV, A, LD, dY, LDV, NP, d, NP, 1.f, 0.f, 0, 1);         // dW2 = A^T dY
  __syncthreads();
  mm<false, true, false>(dA, LD, dY, LDV, W2, LDV, NP, NP, d, 1.f);         // dA  = dY W2^T
  mm<false, true, false>(G, LD, dW2, LDV, W1, LDV, NP, NP, d, 1.f, 0.f, 0, 1);         // dZ  = dW2 W1^T
  mm<true, false, false>(dW1, LDV, Zl, LD, dW2, LDV, NP, d, NP, 1.f, 0.f, 0, 2);       // dW1 = Z^T dW2
  __syncthreads();
  mm<false, true, true>(dA, LD, dW1, LDV, V1, LDV, NP, NP, d, 1.f);         // dA += dW1 V1^T
  mm<true, false, false>(W2, LDV, A, LD, dW1, LDV, NP, d, NP, 1.f, 0.f, 0, 1);         // dV1 = A^T dW1 (W2 is free)
  __syncthreads();
  {
    float* dst = p.dv + (size_t)blockIdx.x * (N - 1) * d;
    const float* cw = p.conv_w ? p.conv_w + (size_t)(blockIdx.x % p.heads) * p.taps : nullptr;
    const int half = p.taps >> 1, n = N - 1;
    for (int idx = threadIdx.x; idx < n * d; idx += kNyThreads) {
      const int i = idx / d, c = idx - i * d;
      float acc = W2[(i + 1) * LDV + c];
      if (cw) {                                     // transpose of the convolution: dv[i] += sum_t w[t] dy[i - t + half]
        const int t0 = max(0, i + half - n + 1), t1 = min(p.taps, i + half + 1);
        for (int t = t0; t < t1; ++t) acc = fmaf(__ldg(cw + t), dY[(i - t + half + 1) * LDV + c], acc);
      }
      dst[idx] = acc;
    }
    if (cw) {                                       // dw[t] = sum_{i,c} dy[i][c] v[i + t - half][c], one warp per tap
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (int t = warp; t < p.taps; t += kNyThreads / 32) {
        const int i0 = max(0, half - t), i1 = min(n, n + half - t);
        float acc = 0.f;
        for (int idx = i0 * d + lane; idx < i1 * d; idx += 32) {
          const int i = idx / d, c = idx - i * d;
          acc = fmaf(dY[(i + 1) * LDV + c], V1[(i + t - half + 1) * LDV + c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) p.dconv[(size_t)blockIdx.x * p.taps + t] = acc;
      }
    }
  }
  __syncthreads();                                // the vectors are dead: the scratch matrices may be overwritten
  // ---- the iteration, backwards ----
  for (int k = p.iters - 1; k >= 0; --k) {
    const float* Z = Zs + (size_t)k * MS;
    ns_terms(A, Z, AZ, T1, T2, T3, N, NP, LD);    // AZ, T1 = 7I - AZ, T2 = 15I - AZ T1, T3 = 13I - AZ T2
    mm<false, true, false>(Gn, LD, G, LD, T3, LD, NP, NP, NP, 0.25f);       // dZ_k  = 1/4 G T3^T
    mm<true, false, false>(X1, LD, Z, LD, G, LD, NP, NP, NP, 0.25f, 0.f, 0, 1);        // dT3   = 1/4 Z^T G
    __syncthreads();
    mm<false, true, false>(dAZ, LD, X1, LD, T2, LD, NP, NP, NP, -1.f);      // dAZ   = -dT3 T2^T
    mm<true, false, false>(X2, LD, AZ, LD, X1, LD, NP, NP, NP, -1.f, 0.f, 0, 1);       // dT2   = -AZ^T dT3
    __syncthreads();
    mm<false, true, true>(dAZ, LD, X2, LD, T1, LD, NP, NP, NP, -1.f);       // dAZ  -= dT2 T1^T
    mm<true, false, false>(X1, LD, AZ, LD, X2, LD, NP, NP, NP, -1.f, 0.f, 0, 1);       // dT1   = -AZ^T dT2
    __syncthreads();
    for (int idx = threadIdx.x; idx < NP * NP; idx += kNyThreads) {
      const int i = idx / NP, j = idx - i * NP;
      dAZ[i * LD + j] -= X1[i * LD + j];                                    // AZ enters T1 with a minus sign
    }
    __syncthreads();
    mm<false, true, true>(dA, LD, dAZ, LD, Z, LD, NP, NP, NP, 1.f);         // dA   += dAZ Z^T
    mm<true, false, true>(Gn, LD, A, LD, dAZ, LD, NP, NP, NP, 1.f, 0.f, 0, 1);         // dZ_k += A^T dAZ
    __syncthreads();
    float* t = G; G = Gn; Gn = t;
  }
  // ---- Z_0 = s A^T ----
  float part = 0.f;
  float* dst = p.dmat + (size_t)blockIdx.x * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += kNyThreads) {
    const int i = idx / N, j = idx - i * N;
    const float g = G[j * LD + i];                // dZ_0[j][i] multiplies A[i][j]
    dst[idx] = dA[i * LD + j] + s * g;
    part = fmaf(g, A[i * LD + j], part);
  }
  part = warp_sum(part);
  __shared__ float s_part[kNyThreads / 32];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kNyThreads / 32; ++w) t += s_part[w];
    p.dscale[blockIdx.x] = t;
  }
}

size_t ny_fwd_smem(int N, int d) {
  const int NP = (N + 3) & ~3;
  return ((size_t)7 * NP * (NP + 1) + (size_t)3 * NP * (d + 1)) * sizeof(float);
}
size_t ny_bwd_smem(int N, int d, int iters) {
  const int NP = (N + 3) & ~3;
  const size_t MS = (size_t)NP * (NP + 1), VS = (size_t)NP * (d + 1);
  return ((size_t)(iters + 12) * MS + (6 * VS <= 7 * MS ? 0 : 6 * VS)) * sizeof(float);
}

int ny_check(int BH, int N, int d, int iters, const char* who) {
  if (BH <= 0 || N < 2 || iters < 0) IMP_FAIL(IMP_ERR_ARG, "%s: sizes must be positive (BH %d, N %d, iters %d)", who, BH, N, iters);
  if (((N + 3) & ~3) > kNyMaxN) IMP_FAIL(IMP_ERR_ARG, "%s: at most %d tokens per matrix (got N = %d)", who, kNyMaxN - 1, N);
  if (d != 32 && d != 64) IMP_FAIL(IMP_ERR_ARG, "%s: head dim must be 32 or 64 (got %d)", who, d);
  if (iters > 8) IMP_FAIL(IMP_ERR_ARG, "%s: at most 8 iterations (got %d)", who, iters);
  return IMP_OK;
}

// ------------------------------------------------------------------------------------------
// The reduced matrix itself (token_tail.nystrom_short): from q (already scaled) and k of the n real tokens,
//   s = q k^T;  soft-max over {p zero logits of the padded tokens, s_i*}:  c_i = e^{-max}/Z_i,  D_ij = e^{s_ij-max}/Z_i
//   M = [[p/m, sqrt(p)/m 1^T], [sqrt(p) c, D]]   ((n+1) x (n+1)),
// plus, per matrix, the largest row sum p c_i + sum_j D_ij and the largest column sum of the FULL m x m matrix
// (p/m + sum_i c_i for a padded column, p/m + sum_i D_ij for a token): the caller takes their maxima over the batch for
// the pseudo-inverse's initial scale (ops/utils.py:119-121) and autograd routes the gradient of that scale back into
// the arg-max row / column here.  One CTA per (slide, head); ~35 forward and ~60 backward library launches per layer
// call become one each.
// ------------------------------------------------------------------------------------------
constexpr int kNbThreads = 256;
struct NyBuildParams {
  const float* q;          // (BH, n, d) scaled queries
  const float* k;          // (BH, n, d)
  const float* dmat;       // (BH, n+1, n+1)      backward
  const float* drow;       // (BH)                backward: gradient of rowmax
  const float* dcol;       // (BH)                backward: gradient of colmax
  float* mat;              // (BH, n+1, n+1)      forward
  float* rowmax;           // (BH)
  float* colmax;           // (BH)
  float* dq;               // (BH, n, d)          backward
  float* dk;
  int n, d;
  float p, inv_m;          // number of padded tokens, 1 / landmarks
};

// shared: Q, K [n][d+1], S [n][n+1] (s, then D), c [n], rsum [n], csum [n+1], arg[2]
__device__ __forceinline__ void ny_build_common(const NyBuildParams& p, float* Q, float* K, float* S, float* c, float* rsum,
                                                float* csum, int* arg) {
  const int n = p.n, d = p.d, LQ = d + 1, LS = n + 1;
  const float* q = p.q + (size_t)blockIdx.x * n * d;
  const float* k = p.k + (size_t)blockIdx.x * n * d;
  for (int idx = threadIdx.x; idx < n * d; idx += kNbThreads) {
    const int i = idx / d, e = idx - i * d;
    Q[i * LQ + e] = __ldg(q + idx);
    K[i * LQ + e] = __ldg(k + idx);
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < n * n; idx += kNbThreads) {
    const int i = idx / n, j = idx - i * n;
    float a = 0.f;
    for (int e = 0; e < d; ++e) a = fmaf(Q[i * LQ + e], K[j * LQ + e], a);
    S[i * LS + j] = a;
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {                       // thread = row: soft-max with p zero logits in front
    const int i = threadIdx.x;
    float mx = 0.f;
    for (int j = 0; j < n; ++j) mx = fmaxf(mx, S[i * LS + j]);
    float z = p.p * __expf(-mx);
    for (int j = 0; j < n; ++j) { const float e = __expf(S[i * LS + j] - mx); S[i * LS + j] = e; z += e; }
    const float iz = 1.f / z;
    float rs = 0.f;
    for (int j = 0; j < n; ++j) { const float v = S[i * LS + j] * iz; S[i * LS + j] = v; rs += v; }
    c[i] = __expf(-mx) * iz;
    rsum[i] = p.p * c[i] + rs;
  }
  __syncthreads();
  if ((int)threadIdx.x <= n) {                      // thread = column of the full matrix (0 = a padded token)
    const int j = threadIdx.x;
    float cs = p.p * p.inv_m;
    if (j == 0) { for (int i = 0; i < n; ++i) cs += c[i]; }
    else { for (int i = 0; i < n; ++i) cs += S[i * LS + j - 1]; }
    csum[j] = cs;
  }
  __syncthreads();
  if (threadIdx.x == 0) {                           // first maximum wins, as in a left-to-right scan
    int ra = 0, ca = 0;
    for (int i = 1; i < n; ++i) if (rsum[i] > rsum[ra]) ra = i;
    for (int j = 1; j <= n; ++j) if (csum[j] > csum[ca]) ca = j;
    arg[0] = ra; arg[1] = ca;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kNbThreads) nystrom_build_fwd_kernel(const NyBuildParams p) {
  extern __shared__ float nb_smem[];
  const int n = p.n, d = p.d, LQ = d + 1, LS = n + 1, N = n + 1;
  float* Q = nb_smem;
  float* K = Q + n * LQ;
  float* S = K + n * LQ;
  float* c = S + n * LS;
  float* rsum = c + n;
  float* csum = rsum + n;
  int* arg = reinterpret_cast<int*>(csum + n + 1);
  ny_build_common(p, Q, K, S, c, rsum, csum, arg);
  float* mat = p.mat + (size_t)blockIdx.x * N * N;
  const float rp = sqrtf(p.p);
  for (int idx = threadIdx.x; idx < N * N; idx += kNbThreads) {
    const int i = idx / N, j = idx - i * N;
    float v;
    if (i == 0) v = (j == 0 ? p.p : rp) * p.inv_m;
    else v = j == 0 ? rp * c[i - 1] : S[(i - 1) * LS + j - 1];
    mat[idx] = v;
  }
  if (threadIdx.x == 0) {
    p.rowmax[blockIdx.x] = rsum[arg[0]];
    p.colmax[blockIdx.x] = csum[arg[1]];
  }
}

__global__ void __launch_bounds__(kNbThreads) nystrom_build_bwd_kernel(const NyBuildParams p) {
  extern __shared__ float nb_smem[];
  const int n = p.n, d = p.d, LQ = d + 1, LS = n + 1, N = n + 1;
  float* Q = nb_smem;
  float* K = Q + n * LQ;
  float* S = K + n * LQ;                            // D
  float* c = S + n * LS;
  float* rsum = c + n;
  float* csum = rsum + n;
  int* arg = reinterpret_cast<int*>(csum + n + 1);
  float* G = reinterpret_cast<float*>(arg + 2);     // dD, then ds   [n][n+1]
  float* dc = G + n * LS;                           // [n]
  ny_build_common(p, Q, K, S, c, rsum, csum, arg);
  const float* dmat = p.dmat + (size_t)blockIdx.x * N * N;
  const float drow = __ldg(p.drow + blockIdx.x), dcol = __ldg(p.dcol + blockIdx.x);
  const int ra = arg[0], ca = arg[1];
  const float rp = sqrtf(p.p);
  for (int idx = threadIdx.x; idx < n * n; idx += kNbThreads) {
    const int i = idx / n, j = idx - i * n;
    G[i * LS + j] = __ldg(dmat + (size_t)(i + 1) * N + j + 1) + (i == ra ? drow : 0.f) + (j + 1 == ca ? dcol : 0.f);
  }
  if ((int)threadIdx.x < n) {
    const int i = threadIdx.x;
    dc[i] = rp * __ldg(dmat + (size_t)(i + 1) * N) + (i == ra ? p.p * drow : 0.f) + (ca == 0 ? dcol : 0.f);
  }
  __syncthreads();
  if ((int)threadIdx.x < n) {                       // ds_ij = D_ij (dD_ij - sum_k dD_ik D_ik - dc_i c_i)
    const int i = threadIdx.x;
    float inner = dc[i] * c[i];
    for (int j = 0; j < n; ++j) inner = fmaf(G[i * LS + j], S[i * LS + j], inner);
    for (int j = 0; j < n; ++j) G[i * LS + j] = S[i * LS + j] * (G[i * LS + j] - inner);
  }
  __syncthreads();
  float* dq = p.dq + (size_t)blockIdx.x * n * d;
  float* dk = p.dk + (size_t)blockIdx.x * n * d;
  for (int idx = threadIdx.x; idx < n * d; idx += kNbThreads) {
    const int i = idx / d, e = idx - i * d;
    float a = 0.f, b = 0.f;
    for (int j = 0; j < n; ++j) {
      a = fmaf(G[i * LS + j], K[j * LQ + e], a);    // dq_i = sum_j ds_ij k_j
      b = fmaf(G[j * LS + i], Q[j * LQ + e], b);    // dk_i = sum_j ds_ji q_j
    }
    dq[idx] = a;
    dk[idx] = b;
  }
}

static size_t ny_build_smem(int n, int d, bool bwd) {
  size_t f = (size_t)2 * n * (d + 1) + (size_t)n * (n + 1) + n + n + (n + 1) + 2;
  if (bwd) f += (size_t)n * (n + 1) + n;
  return f * sizeof(float);
}
static int ny_build_check(int BH, int n, int d, int landmarks, const char* who) {
  if (BH <= 0 || n < 1 || n > kNyMaxN - 1) IMP_FAIL(IMP_ERR_ARG, "%s: 1..%d tokens per matrix (got %d; %d matrices)", who, kNyMaxN - 1, n, BH);
  if (d < 1 || d > 128) IMP_FAIL(IMP_ERR_ARG, "%s: head dim %d out of [1,128]", who, d);
  if (landmarks <= n) IMP_FAIL(IMP_ERR_ARG, "%s: %d landmarks need more than the %d tokens", who, landmarks, n);
  return IMP_OK;
}

}  // namespace

int launch_nystrom_build_fwd(const float* q, const float* k, int BH, int n, int d, int landmarks, float* mat,
                             float* rowmax, float* colmax, cudaStream_t st) {
  { const int rc = ny_build_check(BH, n, d, landmarks, "nystrom_build_fwd"); if (rc) return rc; }
  NyBuildParams p{};
  p.q = q; p.k = k; p.mat = mat; p.rowmax = rowmax; p.colmax = colmax; p.n = n; p.d = d;
  p.p = (float)(landmarks - n); p.inv_m = 1.f / (float)landmarks;
  const size_t smem = ny_build_smem(n, d, false);
  { const int rc = imp_ensure_smem((const void*)nystrom_build_fwd_kernel, 96 * 1024); if (rc) return rc; }
  IMP_LAUNCH("nystrom_build_fwd", st, nystrom_build_fwd_kernel<<<BH, kNbThreads, smem, st>>>(p));
  return IMP_OK;
}

int launch_nystrom_build_bwd(const float* q, const float* k, const float* dmat, const float* drow, const float* dcol,
                             int BH, int n, int d, int landmarks, float* dq, float* dk, cudaStream_t st) {
  { const int rc = ny_build_check(BH, n, d, landmarks, "nystrom_build_bwd"); if (rc) return rc; }
  NyBuildParams p{};
  p.q = q; p.k = k; p.dmat = dmat; p.drow = drow; p.dcol = dcol; p.dq = dq; p.dk = dk; p.n = n; p.d = d;
  p.p = (float)(landmarks - n); p.inv_m = 1.f / (float)landmarks;
  const size_t smem = ny_build_smem(n, d, true);
  { const int rc = imp_ensure_smem((const void*)nystrom_build_bwd_kernel, 96 * 1024); if (rc) return rc; }
  IMP_LAUNCH("nystrom_build_bwd", st, nystrom_build_bwd_kernel<<<BH, kNbThreads, smem, st>>>(p));
  return IMP_OK;
}

size_t nystrom_core_saved_floats(int N, int iters) {
  const int NP = (N + 3) & ~3;
  return (size_t)(iters + 1) * NP * (NP + 1);
}

static int ny_check_conv(const float* conv_w, int heads, int taps, int BH, const char* who) {
  if (!conv_w) return IMP_OK;
  if (heads <= 0 || BH % heads) IMP_FAIL(IMP_ERR_ARG, "%s: %d matrices are not a multiple of %d heads", who, BH, heads);
  if (taps <= 0 || !(taps & 1) || taps > 129) IMP_FAIL(IMP_ERR_ARG, "%s: the residual convolution needs an odd number of taps <= 129 (got %d)", who, taps);
  return IMP_OK;
}

int launch_nystrom_core_fwd(const float* mat, const float* inv_scale, const float* v, const float* conv_w, int heads,
                            int taps, int BH, int N, int d, int iters, float* y, float* zs, cudaStream_t st) {
  { const int rc = ny_check(BH, N, d, iters, "nystrom_core_fwd"); if (rc) return rc; }
  { const int rc = ny_check_conv(conv_w, heads, taps, BH, "nystrom_core_fwd"); if (rc) return rc; }
  NyParams p{};
  p.conv_w = conv_w; p.heads = conv_w ? heads : 1; p.taps = conv_w ? taps : 1;
  p.mat = mat; p.inv_scale = inv_scale; p.v = v; p.y = y; p.zs = zs; p.N = N; p.d = d; p.iters = iters;
  const size_t smem = ny_fwd_smem(N, d);
  { const int rc = imp_ensure_smem((const void*)nystrom_core_fwd_kernel, 227 * 1024 - 1024); if (rc) return rc; }
  IMP_LAUNCH("nystrom_core_fwd", st, nystrom_core_fwd_kernel<<<BH, kNyThreads, smem, st>>>(p));
  return IMP_OK;
}

int launch_nystrom_core_bwd(const float* mat, const float* inv_scale, const float* v, const float* dy, const float* zs,
                            const float* conv_w, int heads, int taps, int BH, int N, int d, int iters, float* dmat,
                            float* dscale, float* dv, float* dconv, cudaStream_t st) {
  { const int rc = ny_check(BH, N, d, iters, "nystrom_core_bwd"); if (rc) return rc; }
  { const int rc = ny_check_conv(conv_w, heads, taps, BH, "nystrom_core_bwd"); if (rc) return rc; }
  if (conv_w && !dconv) IMP_FAIL(IMP_ERR_ARG, "nystrom_core_bwd: conv_w without dconv");
  NyParams p{};
  p.conv_w = conv_w; p.dconv = dconv; p.heads = conv_w ? heads : 1; p.taps = conv_w ? taps : 1;
  p.mat = mat; p.inv_scale = inv_scale; p.v = v; p.dy = dy; p.dmat = dmat; p.dscale = dscale; p.dv = dv;
  p.zs = const_cast<float*>(zs);
  p.N = N; p.d = d; p.iters = iters;
  const size_t smem = ny_bwd_smem(N, d, iters);
  if (smem > 227 * 1024 - 1024) IMP_FAIL(IMP_ERR_ARG, "nystrom_core_bwd: N = %d with %d iterations needs %zu bytes of shared memory", N, iters, smem);
  { const int rc = imp_ensure_smem((const void*)nystrom_core_bwd_kernel, 227 * 1024 - 1024); if (rc) return rc; }
  IMP_LAUNCH("nystrom_core_bwd", st, nystrom_core_bwd_kernel<<<BH, kNyThreads, smem, st>>>(p));
  return IMP_OK;
}
