/* libimp_sm100.so -- C-ABI of the B200-native IMP prototype-fusion hot path.
 *
 * The reference (helenypzhang/Interpretable-Multimodal-Prototyping) is pure Python/PyTorch and has
 * no FFI of its own; each entry point below names the reference arithmetic it
 * replaces (file:line under the reference root).  INTEGRATION.md shows the ctypes stub a maintainer
 * of the reference would add.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated; `stream` is a cudaStream_t passed
 * as void*; functions return 0 on success, non-zero on failure with a message available from
 * imp_last_error() (thread-local).  Nothing allocates: workspaces are sized by the matching
 * *_workspace_bytes() call and owned by the caller.  Calls are asynchronous on `stream` and
 * re-entrant across distinct streams.  bf16 tensors are passed as void*.
 */
#ifndef IMP_HOTPATH_H
#define IMP_HOTPATH_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IMP_ABI_VERSION 1

const char* imp_last_error(void);
int imp_abi_version(void);

/* Launch accounting (measurement only).  imp_launch_count: kernels launched by this library so far.
 * imp_profile_enable(1): bracket every launch with CUDA events on its stream; imp_profile_collect
 * synchronises the device and returns (kernel name, milliseconds) for the launches since the last
 * collect (HOST arrays, up to max_records; returns the count, -1 on a CUDA error). */
long long imp_launch_count(void);
int imp_profile_enable(int on);
int imp_profile_collect(const char** names, float* ms, int max_records);

/* A1  path_net: h = Dropout_p(ReLU(x W1^T + b1))        medmm/modeling/models/umeml_gan.py:266-268,410
 * x (rows,in_features) bf16 row-major, w1 (256,in_features) bf16, b1 (256) fp32 -> h (rows,256) bf16.
 * The dropout keep-mask is a stateless hash of (seed,row,col); the same seed regenerates it. */
int imp_pathnet_fwd(const void* x, const void* w1, const float* b1, void* h, int rows, int in_features,
                    float p_drop, unsigned seed, void* stream);

/* Optional device word XOR-ed into the seed of every dropout kernel (path_net, omic encoders) at run time;
 * NULL (the default) disables it.  A captured CUDA graph bakes `seed` arguments in; advancing this word inside
 * the graph gives every replay a fresh keep-mask (umeml_gan.py:267 draws a new mask per step). */
int imp_set_seed_offset(const unsigned* device_word);

/* backward of A1 wrt W1: dw1 (256,in_features) fp32 (+)= dz^T x, dz (rows,256) bf16.   autograd of :266 */
size_t imp_pathnet_dw_workspace_bytes(int in_features);
int imp_pathnet_dw(const void* dz, const void* x, float* dw1, void* workspace, int rows, int in_features,
                   int accumulate, void* stream);

/* A2/A3  softmax pooling of the patches of each bag into P prototype tokens (keys = values = h):
 *   medmm/modeling/models/umeml_gan.py:65-80,425-434 ; medmm/modeling/ops/attention.py:345-533
 * With the folded query qt = ((c Wq^T + bq)/16) Wk (P,256) [attention.py:368,382,432,509]:
 *   S = qt h^T ; a = softmax over patches [attention.py:527] ; pooled = a h [attention.py:530]
 * h (total_rows,256) bf16, bags packed back to back; cu_seqlens (n_bags+1) int32 row offsets;
 * max_len >= the longest bag (host-side bound used for the launch grid; no device sync);
 * qt (n_bags or 1, P, 256) fp32 with qt_bag_stride elements between bags (0 = shared queries);
 * pooled (n_bags,P,256) fp32, lse (n_bags,P) fp32 = log sum_n exp(S).  1 <= P <= 64. */
size_t imp_pool_fwd_workspace_bytes(int n_bags, int max_len, int n_proto);
int imp_pool_fwd(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                 const float* qt, long long qt_bag_stride, int n_proto, void* workspace, float* pooled,
                 float* lse, void* stream);

/* backward of n_blocks (1 or 2) stacked pooling blocks that share h (autograd of the lines above).
 * Per block k (HOST arrays of n_blocks device pointers): qt[k], dpooled[k] (n_bags,P,256), lse[k],
 * delta[k] (n_bags,P) with delta = rowsum(dpooled * pooled).
 * dq (n_bags,P,256) fp32 <- gradient wrt qt[dq_block].
 * If dz != NULL also: dz (total_rows,256) bf16 <- dL/dh summed over the blocks when relu_mask = 0,
 * or keep_scale * [h>0] * dL/dh (the gradient at the pre-activation of path_net, umeml_gan.py:266-268)
 * when relu_mask = 1; and db1 (256) (+)= column sums of dz. */
size_t imp_pool_bwd_workspace_bytes(int n_bags, int max_len, int n_proto);
int imp_pool_bwd(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                 int n_blocks, const float* const* qt, const long long* qt_bag_stride,
                 const float* const* dpooled, const float* const* lse, const float* const* delta,
                 int n_proto, int dq_block, int relu_mask, float keep_scale, void* workspace, float* dq,
                 void* dz,
                 float* db1, int db_accumulate, void* stream);

/* A0  sentinel strip: medmm/modeling/models/umeml_gan.py:401-410 (+ pad value data_manager.py:387).
 * img (n_bags,n_pad,dim) fp32 in the reference batch layout; lengths[b] = index of the first row
 * that holds an element equal to `sentinel` (-10000), or n_pad when there is none (the reference
 * leaves that case undefined).  If cu_seqlens != NULL it receives the (n_bags+1) exclusive prefix
 * sums, computed on the device (replaces the reference's per-slide .item() host sync). */
int imp_bag_lengths(const float* img, int n_bags, int n_pad, int dim, float sentinel, int* lengths,
                    int* cu_seqlens, void* stream);
/* valid rows of img -> x_packed (cu_seqlens[n_bags], dim) bf16, bag b at rows [cu[b], cu[b+1]).
 * x_packed must have room for n_bags*n_pad rows unless the caller knows the total. */
int imp_pack_bags(const float* img, int n_bags, int n_pad, int dim, const int* cu_seqlens, void* x_packed,
                  void* stream);
/* fp32 -> bf16 round-to-nearest-even of n elements (n % 4 == 0); weights and packed features. */
int imp_cast_bf16(const float* src, void* dst, size_t n, void* stream);

/* A4-A6  modularity loss and its gradient: medmm/modeling/ops/utils.py:178-228 (cluster_assignment_matrix,
 * get_modularity_matrix_and_edge, compute_modularity), call sites umeml_gan.py:516-529.
 * h (total_rows,256) bf16 packed bags (treated as detached, utils.py:208).  chat (n_bags, n_tok1+n_tok2, 256)
 * fp32: the token matrices ALREADY normalised across tokens per feature (the F.normalize(dim=1) on c.T,
 * utils.py:180,214, stays in the caller so autograd handles it); tokens [0,n_tok1) and
 * [n_tok1,n_tok1+n_tok2) are two independent groups evaluated against the same patch graph
 * (prototype tokens and omic tokens).  n_tok1 in [1,32], n_tok2 in [0,8].
 * loss (n_bags,2) fp32 <- -100 tr((W/e) delta) per group (utils.py:222-228; 0 for an empty group);
 * dchat (n_bags, n_tok1+n_tok2, 256) fp32 <- d loss_group / d chat. */
size_t imp_modularity_workspace_bytes(int total_rows, int n_bags, int n_tok1, int n_tok2);
/* Host-only query (no launch): the column split the pair sweep uses for bags of at most max_len patches of which this
 * call owns runs of at most own_rows rows (own_rows = max_len unless the bag is sharded): grid = (ceil(own_rows/128),
 * *nsplit, n_bags) CTAs of *tiles_per_split 64-column tiles each. */
int imp_modularity_sweep_plan(int own_rows, int max_len, int n_bags, int* nsplit, int* tiles_per_split);
int imp_modularity(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                   const float* chat, int n_tok1, int n_tok2, float temp, void* workspace, float* loss,
                   float* dchat, void* stream);

/* The same computation for ONE bag whose patches are sharded over ranks by rows (multi-GPU giant bag, SURVEY.md
 * 8(e); no reference counterpart).  Every rank allocates the full workspace and runs
 *   imp_modularity_prepare  on its rows [row_offset, row_offset+local_rows) (row_offset % 64 == 0; h_local points
 *                           at its first row): zeroes the accumulators, writes its part of the sections below;
 *   collectives by the caller: exchange the row ranges of sections 0 (xh, 512 B per row) and 1 (fixed-point
 *                           assignments, n_slot*256 B per 64-row tile, n_slot = sizes[1]/tiles), SUM-all-reduce
 *                           section 2 (column sums, n_bags*256 floats), MAX-all-reduce section 3 (sign flag, 1 int);
 *   imp_modularity_execute  degrees over all rows, pair sweep and token gradients of its rows against all columns:
 *                           loss (n_bags,2) and dchat are PARTIAL sums the caller SUM-all-reduces.
 * imp_modularity_sections returns byte offsets / sizes of the four sections inside the workspace (HOST arrays of 4).
 * With row_offset = 0 and local_rows = total_rows the two calls are exactly imp_modularity (any n_bags). */
int imp_modularity_sections(int total_rows, int n_bags, int n_tok1, int n_tok2, size_t* offsets, size_t* sizes);
int imp_modularity_prepare(const void* h_local, int local_rows, int row_offset, int total_rows, const int* cu_seqlens,
                           int n_bags, const float* chat, int n_tok1, int n_tok2, void* workspace, void* stream);
int imp_modularity_execute(const void* h_local, int local_rows, int row_offset, int total_rows, const int* cu_seqlens,
                           int n_bags, int max_len, int n_tok1, int n_tok2, float temp, void* workspace, float* loss,
                           float* dchat, void* stream);

/* A7  per-pathway omic encoders: medmm/modeling/models/umeml_gan.py:274-283,413-419, with the
 * feature-level imputation of :391-392 fused into the gather.
 * x_omic (batch,n_genes) fp32; insample_mask (batch,n_genes) int32 or NULL (1 = gene missing, its
 * value is replaced by omic_means[gene]); gene_index: concatenated column indices of the groups,
 * group k at [group_offsets[k], group_offsets[k+1]) (group_offsets is a HOST array of n_groups+1 ints);
 * weights[k] (256, G_k), biases[k] (256) fp32 (HOST arrays of device pointers, n_groups <= 8).
 * out (batch, n_groups, 256) fp32 = Dropout_p(ReLU(x[:, idx_k] W_k^T + b_k)). */
int imp_omic_encode_fwd(const float* x_omic, const int* insample_mask, const float* omic_means,
                        const int* gene_index, const int* group_offsets, int n_groups,
                        const float* const* weights, const float* const* biases, int batch, int n_genes,
                        float p_drop, unsigned seed, float* out, void* stream);
/* backward of A7 wrt the encoder weights/biases (inputs carry no gradient): dweights[k] (256,G_k),
 * dbiases[k] (256) (+)= ...; `out` is the forward output (its zeros are the ReLU/dropout mask). */
int imp_omic_encode_bwd(const float* x_omic, const int* insample_mask, const float* omic_means,
                        const int* gene_index, const int* group_offsets, int n_groups, int batch, int n_genes,
                        float p_drop, const float* out, const float* dout, float* const* dweights,
                        float* const* dbiases, int accumulate, void* stream);
/* A8  missing-omics handling of the omic tokens: umeml_gan.py:500-511.
 *   h = where(without_omic[b] == 1, gen, h_omic)             (:503-505; without_omic (batch) int32 or NULL)
 *   r = sum(insample_mask)/mask_numel over the whole batch   (:509; NULL mask -> r = 0)
 *   out = (1-r) h + r gen                                    (:510-511)
 * The reference's host-side "if sum(...) > 0" tests become device no-ops; scratch = 1 float;
 * ratio_out (1 float, may be NULL) receives r.  h_omic, gen, out: (batch, per_sample) fp32. */
int imp_omic_blend(const float* h_omic, const float* h_omic_gen, const int* without_omic,
                   const int* insample_mask, long long mask_numel, int batch, int per_sample,
                   float* scratch, float* out, float* ratio_out, void* stream);

/* A9  k-means prototype assignment (NEW: the reference has no k-means; distance formula of
 * medmm/metrics/distance.py:46-61 followed by argmin, first index on ties).
 * x (n,dim) fp32, centroids (k,dim) fp32, k <= 64, dim % 32 == 0 -> assign (n) int32;
 * best_dist (n) fp32 (may be NULL) <- the winning squared distance. */
int imp_kmeans_assign(const float* x, const float* centroids, int n, int dim, int k, int* assign,
                      float* best_dist, void* stream);
/* Lloyd update: sums (k,dim) fp32 += x rows per assigned centroid, counts (k) int32 += 1 (caller zeroes). */
int imp_kmeans_update(const float* x, const int* assign, int n, int dim, int k, float* sums, int* counts,
                      void* stream);

/* Multi-GPU giant bag: log-sum-exp merge of per-shard pooling results.  Each of the
 * n_parts shards of a bag produced (pooled_r, lse_r) with imp_pool_fwd over its patches;
 * part_pooled (n_bags, n_parts, P, 256), part_lse (n_bags, n_parts, P) (an empty shard has lse = -inf)
 * -> pooled (n_bags,P,256), lse (n_bags,P) of the whole bag.  scratch: n_bags*n_parts*2*P floats. */
int imp_lse_merge(const float* part_pooled, const float* part_lse, int n_bags, int n_parts, int n_proto,
                  float* pooled, float* lse, float* scratch, void* stream);

/* N1  token-level tail: the core of NystromAttention for sequences shorter than the landmark count
 * (medmm/modeling/ops/attention.py:105-131 with moore_penrose_iter_pinv, ops/utils.py:116-131), on the reduced
 * matrices M(A) of imp_b200.token_tail.nystrom_short: for each of n_mat (slide, head) pairs
 *   Z_0 = s M^T;  Z_{k+1} = 1/4 Z_k (13 I - M Z_k (15 I - M Z_k (7 I - M Z_k))), k < iters (<= 8);
 *   y = rows 1.. of M (Z (M [0; v]))  +  res_conv(v)     (the depth-wise residual convolution of :129-131 when conv_w != NULL).
 * mat (n_mat, n_dim, n_dim) fp32 with n_dim = tokens + 1 <= 48; inv_scale: ONE device float s (the reference's
 * 1 / (max row-sum * max column-sum) over the whole batch); v, y (n_mat, n_dim - 1, head_dim) fp32, head_dim 32 or 64;
 * conv_w (heads, taps) fp32 or NULL, taps odd, matrix index = slide * heads + head.
 * One CTA per matrix, everything in shared memory.  saved (may be NULL): n_mat * imp_nystrom_core_saved_floats(n_dim,
 * iters) floats that receive the iterates Z_0..Z_iters for the backward call. */
size_t imp_nystrom_core_saved_floats(int n_dim, int iters);
int imp_nystrom_core_fwd(const float* mat, const float* inv_scale, const float* v, const float* conv_w, int heads,
                         int taps, int n_mat, int n_dim, int head_dim, int iters, float* y, float* saved, void* stream);
/* Backward of the above: dy (n_mat, n_dim-1, head_dim) -> dmat (n_mat, n_dim, n_dim), dv like v, dscale (n_mat)
 * partial derivatives wrt s and dconv (n_mat, taps) partial derivatives wrt conv_w (the caller sums them over the
 * slides; dconv may be NULL when conv_w is).  saved: the buffer the forward call filled, or NULL (the iteration is
 * then run again). */
int imp_nystrom_core_bwd(const float* mat, const float* inv_scale, const float* v, const float* dy, const float* saved,
                         const float* conv_w, int heads, int taps, int n_mat, int n_dim, int head_dim, int iters,
                         float* dmat, float* dscale, float* dv, float* dconv, void* stream);

/* The reduced matrix of the short-sequence Nystrom layer itself (imp_b200.token_tail.nystrom_short; the three soft-maxes
 * of ops/attention.py:105-110 coincide there).  q (already multiplied by dim_head^-0.5), k: (n_mat, n_tok, head_dim) fp32,
 * n_tok < landmarks, n_tok <= 47.  With p = landmarks - n_tok zero tokens in front:
 *   mat (n_mat, n_tok+1, n_tok+1) <- [[p/m, sqrt(p)/m 1^T], [sqrt(p) c, D]],  c_i, D_ij = soft-max over {p zeros, (q k^T)_i*};
 *   rowmax, colmax (n_mat) <- the largest row / column sum of the full landmarks x landmarks matrix (the caller takes
 *   their maxima over the batch for the scale s of imp_nystrom_core_fwd: ops/utils.py:119-121). */
int imp_nystrom_build_fwd(const float* q, const float* k, int n_mat, int n_tok, int head_dim, int landmarks, float* mat,
                          float* rowmax, float* colmax, void* stream);
/* Backward: dmat like mat, drowmax / dcolmax (n_mat) -> dq, dk like q, k (the soft-max is recomputed; the gradients of
 * the two maxima go to their arg-max row / column). */
int imp_nystrom_build_bwd(const float* q, const float* k, const float* dmat, const float* drowmax, const float* dcolmax,
                          int n_mat, int n_tok, int head_dim, int landmarks, float* dq, float* dk, void* stream);

#ifdef __cplusplus
}
#endif
#endif
