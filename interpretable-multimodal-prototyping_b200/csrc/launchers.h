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

// modularity.cu
size_t modularity_workspace_bytes(int total_rows, int B, int P1, int P2);
void modularity_sweep_plan(int own_len, int max_len, int B, int* nsplit, int* tiles_per_split);
int launch_modularity(const bf16* h, int total_rows, const int* cu, int B, int max_len, const float* chat, int P1, int P2,
                      float temp, void* workspace, float* loss, float* dchat, cudaStream_t st);
void modularity_workspace_sections(int total_rows, int B, int P1, int P2, size_t* offsets, size_t* sizes);
int launch_modularity_prepare(const bf16* h_local, int local_rows, int row_lo, int total_rows, const int* cu, int B,
                              const float* chat, int P1, int P2, void* workspace, cudaStream_t st);
int launch_modularity_execute(const bf16* h_local, int local_rows, int row_lo, int total_rows, const int* cu, int B,
                              int max_len, int P1, int P2, float temp, void* workspace, float* loss, float* dchat,
                              cudaStream_t st);

// omic.cu
int launch_omic_fwd(const float* x, const int* mask, const float* means, const int* idx, const int* group_offsets,
                    int K, const float* const* w, const float* const* bias, int B, int G, float p_drop, unsigned seed,
                    float* out, cudaStream_t st);
int launch_omic_bwd(const float* x, const int* mask, const float* means, const int* idx, const int* group_offsets,
                    int K, int B, int G, float p_drop, const float* out, const float* dout, float* const* dw,
                    float* const* db, int accumulate, cudaStream_t st);
int launch_omic_blend(const float* h_omic, const float* h_gen, const int* without_omic, const int* insample_mask,
                      long long mask_numel, int B, int per_sample, float* scratch, float* out, float* ratio_out,
                      cudaStream_t st);

// kmeans.cu
int launch_kmeans_assign(const float* x, const float* mu, int N, int D, int K, int* assign, float* best_dist,
                         cudaStream_t st);
int launch_kmeans_update(const float* x, const int* assign, int N, int D, int K, float* sums, int* counts,
                         cudaStream_t st);
int launch_lse_merge(const float* part_pooled, const float* part_lse, int B, int nsplit, int P, float* pooled,
                     float* lse, float* scratch, cudaStream_t st);

// nystrom.cu
size_t nystrom_core_saved_floats(int N, int iters);
int launch_nystrom_core_fwd(const float* mat, const float* inv_scale, const float* v, const float* conv_w, int heads,
                            int taps, int BH, int N, int d, int iters, float* y, float* zs, cudaStream_t st);
int launch_nystrom_core_bwd(const float* mat, const float* inv_scale, const float* v, const float* dy, const float* zs,
                            const float* conv_w, int heads, int taps, int BH, int N, int d, int iters, float* dmat,
                            float* dscale, float* dv, float* dconv, cudaStream_t st);
int launch_nystrom_build_fwd(const float* q, const float* k, int BH, int n, int d, int landmarks, float* mat,
                             float* rowmax, float* colmax, cudaStream_t st);
int launch_nystrom_build_bwd(const float* q, const float* k, const float* dmat, const float* drow, const float* dcol,
                             int BH, int n, int d, int landmarks, float* dq, float* dk, cudaStream_t st);
