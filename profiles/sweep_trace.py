"""Per-warp wait counters of the modularity sweep (profiles/r01_sweep_iterations.md).
Build the library with  IMP_NVCC_EXTRA=-DIMP_SWEEP_TRACE python interpretable-multimodal-prototyping_b200/build.py  first."""
import ctypes, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import imp_b200
from imp_b200 import kernels, _lib
B, N, P = 4, 16384, 32
g = torch.Generator(device="cuda").manual_seed(0)
h = torch.relu(torch.randn(B * N, 256, device="cuda", generator=g)).bfloat16()
cu = torch.arange(0, B + 1, device="cuda", dtype=torch.int32) * N
chat = torch.nn.functional.normalize(torch.randn(B, P + 7, 256, device="cuda", generator=g), dim=-1)
lib = _lib.lib()
fn = lib.imp_debug_sweep_trace
fn.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
kernels.modularity(h, cu, N, chat, P, 7, 0.1)
torch.cuda.synchronize()
fn(None, 1)
t0 = time.time()
kernels.modularity(h, cu, N, chat, P, 7, 0.1)
torch.cuda.synchronize()
print("modularity 4 bags: %.2f ms" % ((time.time() - t0) * 1e3))
out = (ctypes.c_ulonglong * 128)()
fn(out, 0)
print("warp  smsp  wait_tfull%  wait_lfull%  ensure%  cycles/tile  lookahead")
for w in range(16):
    t, l, tot, n, look, e = [out[w * 8 + k] for k in range(6)]
    print("%4d %4d %10.1f %11.1f %9.1f %11.0f %9.2f  blocking issues per owned tile %.2f" % (w, w % 4, 100 * t / tot, 100 * l / tot, 100 * e / tot, tot / max(n, 1), look / max(n, 1), out[w * 8 + 6] / max(n / 16, 1)))
