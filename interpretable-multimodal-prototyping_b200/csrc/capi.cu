// C-ABI of libimp_sm100.so (declared in include/imp_hotpath.h).
// Plain pointers + sizes + a cudaStream_t passed as void*; int status (0 = ok), message via
// imp_last_error().  No allocation, no ownership transfer, no torch types.
#include "common.cuh"
#include "launchers.h"
#include "../../include/imp_hotpath.h"
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

thread_local char g_imp_err[512] = {0};

extern "C" const char* imp_last_error(void) { return g_imp_err; }
extern "C" int imp_abi_version(void) { return IMP_ABI_VERSION; }

// ------------------------------------------------------------------------------------------
// driver entry point for cuTensorMapEncodeTiled, resolved lazily so the library has no link-time
// dependency on libcuda (it must load on a GPU-less build box for the symbol check)
// ------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_fn() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}

int imp_make_tmap_2d(CUtensorMap* out, const void* base, CUtensorMapDataType dt, int elt_bytes, uint64_t inner,
                     uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                     CUtensorMapSwizzle swz) {
  encode_tiled_fn fn = get_encode_fn();
  if (!fn) IMP_FAIL(IMP_ERR_CUDA, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) IMP_FAIL(IMP_ERR_ARG, "tensor base %p not 16-byte aligned", base);
  if (row_stride_bytes % 16 != 0) IMP_FAIL(IMP_ERR_ARG, "row stride %llu not a multiple of 16 B", (unsigned long long)row_stride_bytes);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstr[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  (void)elt_bytes;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IMP_FAIL(IMP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return IMP_OK;
}

// SM count of the CURRENT device (a process may drive several GPUs: cached per device ordinal)
int imp_num_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (!n) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember what was granted per
// (kernel, device) so that a second GPU driven by the same process gets its own opt-in.
int imp_ensure_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  IMP_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  size_t& g = granted[std::make_pair(kernel, dev)];
  if (bytes > g) {
    IMP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    g = bytes;
  }
  return IMP_OK;
}

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// ------------------------------------------------------------------------------------------
// launch accounting: a counter of kernel launches and an optional per-launch CUDA-event timer
// ------------------------------------------------------------------------------------------
#include <atomic>
#include <mutex>
#include <vector>
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
std::atomic<long long> g_launches{0};
std::atomic<int> g_prof_on{0};
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof_recs;       // recorded launches, in order
std::vector<ProfRec> g_prof_pool;       // reusable event pairs
thread_local ProfRec g_prof_cur = {nullptr, nullptr, nullptr};
}  // namespace

void imp_prof_begin(const char* name, cudaStream_t st) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_prof_cur.name = nullptr;
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRec r{name, nullptr, nullptr};
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_pool.empty()) { r = g_prof_pool.back(); g_prof_pool.pop_back(); r.name = name; }
  }
  if (!r.a && (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess)) return;
  cudaEventRecord(r.a, st);
  g_prof_cur = r;
}
void imp_prof_end(cudaStream_t st) {
  if (!g_prof_cur.name) return;
  cudaEventRecord(g_prof_cur.b, st);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_recs.push_back(g_prof_cur);
  g_prof_cur.name = nullptr;
}

extern "C" long long imp_launch_count(void) { return g_launches.load(); }
extern "C" int imp_profile_enable(int on) {
  g_prof_on.store(on ? 1 : 0);
  return IMP_OK;
}
// Synchronises the device, then writes up to `max_records` (name, milliseconds) pairs of the launches
// recorded since the last collect; returns the number written (records beyond max are dropped).
extern "C" int imp_profile_collect(const char** names, float* ms, int max_records) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (ProfRec& r : g_prof_recs) {
    if (n < max_records) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { names[n] = r.name; ms[n] = t; ++n; }
    }
    g_prof_pool.push_back(r);
  }
  g_prof_recs.clear();
  return n;
}

// one seed-offset word per device ordinal (the pointer is device memory of the current device)
static std::atomic<const uint32_t*> g_seed_offset[64];
const uint32_t* imp_seed_offset_ptr() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  return g_seed_offset[dev].load();
}
extern "C" int imp_set_seed_offset(const unsigned* device_word) {
  int dev = 0;
  IMP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) IMP_FAIL(IMP_ERR_ARG, "imp_set_seed_offset: device ordinal %d out of range", dev);
  g_seed_offset[dev].store(device_word);
  return IMP_OK;
}

// ------------------------------------------------------------------------------------------
// A1 path_net
// ------------------------------------------------------------------------------------------
extern "C" int imp_pathnet_fwd(const void* x, const void* w1, const float* b1, void* h, int rows, int in_features,
                               float p_drop, unsigned seed, void* stream) {
  if (!x || !w1 || !b1 || !h) IMP_FAIL(IMP_ERR_ARG, "imp_pathnet_fwd: null pointer");
  return launch_pathnet_fwd((const bf16*)x, (const bf16*)w1, b1, (bf16*)h, rows, in_features, p_drop, seed, ST(stream));
}

extern "C" size_t imp_pathnet_dw_workspace_bytes(int in_features) { return pathnet_dw_workspace_bytes(in_features); }

extern "C" int imp_pathnet_dw(const void* dz, const void* x, float* dw1, void* workspace, int rows, int in_features,
                              int accumulate, void* stream) {
  if (!dz || !x || !dw1 || !workspace) IMP_FAIL(IMP_ERR_ARG, "imp_pathnet_dw: null pointer");
  return launch_pathnet_dw((const bf16*)dz, (const bf16*)x, dw1, (float*)workspace, rows, in_features, accumulate, ST(stream));
}

// ------------------------------------------------------------------------------------------
// A2/A3 softmax pooling into prototype tokens
// ------------------------------------------------------------------------------------------
extern "C" size_t imp_pool_fwd_workspace_bytes(int n_bags, int max_len, int n_proto) {
  return pool_fwd_workspace_bytes(n_bags, max_len, n_proto);
}
extern "C" int imp_pool_fwd(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                            const float* qt, long long qt_bag_stride, int n_proto, void* workspace, float* pooled,
                            float* lse, void* stream) {
  if (!h || !cu_seqlens || !qt || !workspace || !pooled || !lse) IMP_FAIL(IMP_ERR_ARG, "imp_pool_fwd: null pointer");
  return launch_pool_fwd((const bf16*)h, total_rows, cu_seqlens, n_bags, max_len, qt, qt_bag_stride, n_proto,
                         (float*)workspace, pooled, lse, ST(stream));
}
extern "C" size_t imp_pool_bwd_workspace_bytes(int n_bags, int max_len, int n_proto) {
  return pool_bwd_workspace_bytes(n_bags, max_len, n_proto);
}
extern "C" int imp_pool_bwd(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                            int n_blocks, const float* const* qt, const long long* qt_bag_stride,
                            const float* const* dpooled, const float* const* lse, const float* const* delta,
                            int n_proto, int dq_block, int relu_mask, float keep_scale, void* workspace, float* dq, void* dz,
                            float* db1, int db_accumulate, void* stream) {
  if (!h || !cu_seqlens || !qt || !qt_bag_stride || !dpooled || !lse || !delta || !workspace || !dq)
    IMP_FAIL(IMP_ERR_ARG, "imp_pool_bwd: null pointer");
  return launch_pool_bwd((const bf16*)h, total_rows, cu_seqlens, n_bags, max_len, n_blocks, qt, qt_bag_stride, dpooled,
                         lse, delta, n_proto, dq_block, relu_mask, keep_scale, (float*)workspace, dq, (bf16*)dz, db1,
                         db_accumulate, ST(stream));
}

// ------------------------------------------------------------------------------------------
// A0 sentinel strip + wire-format conversion
// ------------------------------------------------------------------------------------------
extern "C" int imp_bag_lengths(const float* img, int n_bags, int n_pad, int dim, float sentinel, int* lengths,
                               int* cu_seqlens, void* stream) {
  if (!img || !lengths) IMP_FAIL(IMP_ERR_ARG, "imp_bag_lengths: null pointer");
  return launch_bag_lengths(img, n_bags, n_pad, dim, sentinel, lengths, cu_seqlens, ST(stream));
}
extern "C" int imp_pack_bags(const float* img, int n_bags, int n_pad, int dim, const int* cu_seqlens, void* x_packed,
                             void* stream) {
  if (!img || !cu_seqlens || !x_packed) IMP_FAIL(IMP_ERR_ARG, "imp_pack_bags: null pointer");
  return launch_pack_bags(img, n_bags, n_pad, dim, cu_seqlens, (bf16*)x_packed, ST(stream));
}
extern "C" int imp_cast_bf16(const float* src, void* dst, size_t n, void* stream) {
  if (!src || !dst) IMP_FAIL(IMP_ERR_ARG, "imp_cast_bf16: null pointer");
  return launch_cast_bf16(src, (bf16*)dst, n, ST(stream));
}

// ------------------------------------------------------------------------------------------
// A4-A6 modularity loss + gradient wrt the normalised tokens
// ------------------------------------------------------------------------------------------
extern "C" size_t imp_modularity_workspace_bytes(int total_rows, int n_bags, int n_tok1, int n_tok2) {
  return modularity_workspace_bytes(total_rows, n_bags, n_tok1, n_tok2);
}
extern "C" int imp_modularity_sweep_plan(int own_rows, int max_len, int n_bags, int* nsplit, int* tiles_per_split) {
  if (!nsplit || !tiles_per_split) IMP_FAIL(IMP_ERR_ARG, "imp_modularity_sweep_plan: null pointer");
  if (own_rows <= 0 || max_len <= 0 || n_bags <= 0) IMP_FAIL(IMP_ERR_ARG, "imp_modularity_sweep_plan: sizes must be positive");
  modularity_sweep_plan(own_rows, max_len, n_bags, nsplit, tiles_per_split);
  return IMP_OK;
}
extern "C" int imp_modularity(const void* h, int total_rows, const int* cu_seqlens, int n_bags, int max_len,
                              const float* chat, int n_tok1, int n_tok2, float temp, void* workspace, float* loss,
                              float* dchat, void* stream) {
  if (!h || !cu_seqlens || !chat || !workspace || !loss || !dchat) IMP_FAIL(IMP_ERR_ARG, "imp_modularity: null pointer");
  return launch_modularity((const bf16*)h, total_rows, cu_seqlens, n_bags, max_len, chat, n_tok1, n_tok2, temp,
                           workspace, loss, dchat, ST(stream));
}

// multi-GPU giant bag: the same computation in two phases around the caller's collectives
extern "C" int imp_modularity_sections(int total_rows, int n_bags, int n_tok1, int n_tok2, size_t* offsets, size_t* sizes) {
  if (!offsets || !sizes) IMP_FAIL(IMP_ERR_ARG, "imp_modularity_sections: null pointer");
  modularity_workspace_sections(total_rows, n_bags, n_tok1, n_tok2, offsets, sizes);
  return IMP_OK;
}
extern "C" int imp_modularity_prepare(const void* h_local, int local_rows, int row_offset, int total_rows,
                                      const int* cu_seqlens, int n_bags, const float* chat, int n_tok1, int n_tok2,
                                      void* workspace, void* stream) {
  if ((!h_local && local_rows > 0) || !cu_seqlens || !chat || !workspace) IMP_FAIL(IMP_ERR_ARG, "imp_modularity_prepare: null pointer");
  return launch_modularity_prepare((const bf16*)h_local, local_rows, row_offset, total_rows, cu_seqlens, n_bags, chat,
                                   n_tok1, n_tok2, workspace, ST(stream));
}
extern "C" int imp_modularity_execute(const void* h_local, int local_rows, int row_offset, int total_rows,
                                      const int* cu_seqlens, int n_bags, int max_len, int n_tok1, int n_tok2, float temp,
                                      void* workspace, float* loss, float* dchat, void* stream) {
  if ((!h_local && local_rows > 0) || !cu_seqlens || !workspace || !loss || !dchat) IMP_FAIL(IMP_ERR_ARG, "imp_modularity_execute: null pointer");
  return launch_modularity_execute((const bf16*)h_local, local_rows, row_offset, total_rows, cu_seqlens, n_bags, max_len,
                                   n_tok1, n_tok2, temp, workspace, loss, dchat, ST(stream));
}

// ------------------------------------------------------------------------------------------
// A7 per-pathway omic encoders, A8 missing-omics handling
// ------------------------------------------------------------------------------------------
extern "C" int imp_omic_encode_fwd(const float* x_omic, const int* insample_mask, const float* omic_means,
                                   const int* gene_index, const int* group_offsets, int n_groups,
                                   const float* const* weights, const float* const* biases, int batch, int n_genes,
                                   float p_drop, unsigned seed, float* out, void* stream) {
  if (!x_omic || !gene_index || !group_offsets || !weights || !biases || !out) IMP_FAIL(IMP_ERR_ARG, "imp_omic_encode_fwd: null pointer");
  return launch_omic_fwd(x_omic, insample_mask, omic_means, gene_index, group_offsets, n_groups, weights, biases, batch,
                         n_genes, p_drop, seed, out, ST(stream));
}
extern "C" int imp_omic_encode_bwd(const float* x_omic, const int* insample_mask, const float* omic_means,
                                   const int* gene_index, const int* group_offsets, int n_groups, int batch, int n_genes,
                                   float p_drop, const float* out, const float* dout, float* const* dweights,
                                   float* const* dbiases, int accumulate, void* stream) {
  if (!x_omic || !gene_index || !group_offsets || !out || !dout || !dweights || !dbiases) IMP_FAIL(IMP_ERR_ARG, "imp_omic_encode_bwd: null pointer");
  return launch_omic_bwd(x_omic, insample_mask, omic_means, gene_index, group_offsets, n_groups, batch, n_genes, p_drop,
                         out, dout, dweights, dbiases, accumulate, ST(stream));
}
extern "C" int imp_omic_blend(const float* h_omic, const float* h_omic_gen, const int* without_omic,
                              const int* insample_mask, long long mask_numel, int batch, int per_sample,
                              float* scratch, float* out, float* ratio_out, void* stream) {
  if (!h_omic || !h_omic_gen || !out) IMP_FAIL(IMP_ERR_ARG, "imp_omic_blend: null pointer");
  return launch_omic_blend(h_omic, h_omic_gen, without_omic, insample_mask, mask_numel, batch, per_sample, scratch, out,
                           ratio_out, ST(stream));
}

// ------------------------------------------------------------------------------------------
// A9 k-means prototype assignment / Lloyd update
// ------------------------------------------------------------------------------------------
extern "C" int imp_kmeans_assign(const float* x, const float* centroids, int n, int dim, int k, int* assign,
                                 float* best_dist, void* stream) {
  if (!x || !centroids || !assign) IMP_FAIL(IMP_ERR_ARG, "imp_kmeans_assign: null pointer");
  return launch_kmeans_assign(x, centroids, n, dim, k, assign, best_dist, ST(stream));
}
extern "C" int imp_kmeans_update(const float* x, const int* assign, int n, int dim, int k, float* sums, int* counts,
                                 void* stream) {
  if (!x || !assign || !sums || !counts) IMP_FAIL(IMP_ERR_ARG, "imp_kmeans_update: null pointer");
  return launch_kmeans_update(x, assign, n, dim, k, sums, counts, ST(stream));
}

extern "C" int imp_lse_merge(const float* part_pooled, const float* part_lse, int n_bags, int n_parts, int n_proto,
                             float* pooled, float* lse, float* scratch, void* stream) {
  if (!part_pooled || !part_lse || !pooled || !lse || !scratch) IMP_FAIL(IMP_ERR_ARG, "imp_lse_merge: null pointer");
  return launch_lse_merge(part_pooled, part_lse, n_bags, n_parts, n_proto, pooled, lse, scratch, ST(stream));
}

// ------------------------------------------------------------------------------------------
// N1 token tail: Nystrom attention core on the reduced matrices
// ------------------------------------------------------------------------------------------
extern "C" size_t imp_nystrom_core_saved_floats(int n_dim, int iters) { return nystrom_core_saved_floats(n_dim, iters); }
extern "C" int imp_nystrom_core_fwd(const float* mat, const float* inv_scale, const float* v, const float* conv_w,
                                    int heads, int taps, int n_mat, int n_dim, int head_dim, int iters, float* y,
                                    float* saved, void* stream) {
  if (!mat || !inv_scale || !v || !y) IMP_FAIL(IMP_ERR_ARG, "imp_nystrom_core_fwd: null pointer");
  return launch_nystrom_core_fwd(mat, inv_scale, v, conv_w, heads, taps, n_mat, n_dim, head_dim, iters, y, saved, ST(stream));
}
extern "C" int imp_nystrom_core_bwd(const float* mat, const float* inv_scale, const float* v, const float* dy,
                                    const float* saved, const float* conv_w, int heads, int taps, int n_mat, int n_dim,
                                    int head_dim, int iters, float* dmat, float* dscale, float* dv, float* dconv,
                                    void* stream) {
  if (!mat || !inv_scale || !v || !dy || !dmat || !dscale || !dv) IMP_FAIL(IMP_ERR_ARG, "imp_nystrom_core_bwd: null pointer");
  return launch_nystrom_core_bwd(mat, inv_scale, v, dy, saved, conv_w, heads, taps, n_mat, n_dim, head_dim, iters, dmat,
                                 dscale, dv, dconv, ST(stream));
}
extern "C" int imp_nystrom_build_fwd(const float* q, const float* k, int n_mat, int n_tok, int head_dim, int landmarks,
                                     float* mat, float* rowmax, float* colmax, void* stream) {
  if (!q || !k || !mat || !rowmax || !colmax) IMP_FAIL(IMP_ERR_ARG, "imp_nystrom_build_fwd: null pointer");
  return launch_nystrom_build_fwd(q, k, n_mat, n_tok, head_dim, landmarks, mat, rowmax, colmax, ST(stream));
}
extern "C" int imp_nystrom_build_bwd(const float* q, const float* k, const float* dmat, const float* drowmax,
                                     const float* dcolmax, int n_mat, int n_tok, int head_dim, int landmarks, float* dq,
                                     float* dk, void* stream) {
  if (!q || !k || !dmat || !drowmax || !dcolmax || !dq || !dk) IMP_FAIL(IMP_ERR_ARG, "imp_nystrom_build_bwd: null pointer");
  return launch_nystrom_build_bwd(q, k, dmat, drowmax, dcolmax, n_mat, n_tok, head_dim, landmarks, dq, dk, ST(stream));
}
