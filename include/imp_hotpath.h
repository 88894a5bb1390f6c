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

/* A1  path_net: h = Dropout_p(ReLU(x W1^T + b1))        medmm/modeling/models/umeml_gan.py:266-268,410
 * x (rows,in_features) bf16 row-major, w1 (256,in_features) bf16, b1 (256) fp32 -> h (rows,256) bf16.
 * The dropout keep-mask is a stateless hash of (seed,row,col); the same seed regenerates it. */
int imp_pathnet_fwd(const void* x, const void* w1, const float* b1, void* h, int rows, int in_features,
                    float p_drop, unsigned seed, void* stream);

/* backward of A1 wrt W1: dw1 (256,in_features) fp32 (+)= dz^T x, dz (rows,256) bf16.   autograd of :266 */
size_t imp_pathnet_dw_workspace_bytes(int in_features);
int imp_pathnet_dw(const void* dz, const void* x, float* dw1, void* workspace, int rows, int in_features,
                   int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif
