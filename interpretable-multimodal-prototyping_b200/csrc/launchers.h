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
