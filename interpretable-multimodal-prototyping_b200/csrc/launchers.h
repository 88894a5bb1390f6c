// Internal host launchers (one per kernel family); the C-ABI in capi.cu forwards to these.
#pragma once
#include "common.cuh"

// pathnet.cu
int launch_pathnet_fwd(const bf16* x, const bf16* w1, const float* b1, bf16* h, int rows, int kdim, float p_drop,
                       uint32_t seed, cudaStream_t st);
size_t pathnet_dw_workspace_bytes(int kin);
int launch_pathnet_dw(const bf16* dz, const bf16* x, float* dw, float* workspace, int rows, int kin, int accumulate,
                      cudaStream_t st);
int launch_sum_partials(const float* partial, float* out, int nsplit, size_t n, float scale, int accumulate,
                        cudaStream_t st);

// pool.cu
size_t pool_fwd_workspace_bytes(int B, int max_len, int P);
int launch_pool_fwd(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* qt,
                    long long qt_stride, int P, float* workspace, float* pooled, float* lse, cudaStream_t st);
size_t pool_bwd_workspace_bytes(int B, int max_len, int P);
int launch_pool_bwd(const bf16* h, int total_rows, const int* cu, int B, int max_len, int nblocks,
                    const float* const* qt, const long long* qt_stride, const float* const* dpool,
                    const float* const* lse, const float* const* delta, int P, int dq_block, int relu_mask,
                    float keep_scale, float* workspace, float* dq, bf16* dz, float* db1, int db_accumulate, cudaStream_t st);

// prep.cu
int launch_bag_lengths(const float* img, int B, int npad, int d, float sentinel, int* lengths, int* cu, cudaStream_t st);
int launch_pack_bags(const float* img, int B, int npad, int d, const int* cu, bf16* out, cudaStream_t st);
int launch_cast_bf16(const float* src, bf16* dst, size_t n, cudaStream_t st);
