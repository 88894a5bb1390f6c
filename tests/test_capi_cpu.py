"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol that
include/imp_hotpath.h declares, and the product path refuses to run without CUDA (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    import imp_b200
    from imp_b200 import _lib
    syms = _lib.header_symbols()
    assert len(syms) >= 20 and "imp_pool_fwd" in syms and "imp_modularity" in syms and "imp_kmeans_assign" in syms
    lib = _lib.lib()
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.imp_abi_version() == 1
    # and nothing exported by the library is missing from the header
    out = os.popen("nm -D --defined-only %s" % _lib.LIB_PATH).read()
    exported = set(re.findall(r" T (imp_[a-z0-9_]+)", out))
    assert exported <= set(syms) | {"imp_make_tmap_2d", "imp_num_sms", "imp_prof_begin", "imp_prof_end"}, exported - set(syms)


def test_header_has_no_torch_types():
    text = open(os.path.join(ROOT, "include", "imp_hotpath.h")).read()
    assert "torch" not in text.lower().replace("pytorch", "") and "at::" not in text and "std::" not in text


def test_argument_errors_are_reported_not_crashing():
    from imp_b200 import _lib
    lib = _lib.lib()
    rc = lib.imp_pathnet_fwd(None, None, None, None, 10, 512, ctypes.c_float(0.0), 0, None)
    assert rc != 0 and b"null" in lib.imp_last_error()
    lib.imp_kmeans_assign.restype = ctypes.c_int
    rc = lib.imp_kmeans_assign(ctypes.c_void_p(16), ctypes.c_void_p(16), 10, 500, 99, ctypes.c_void_p(16), None, None)
    assert rc != 0 and b"K=99" in lib.imp_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from imp_b200 import _lib, kernels, model
    with pytest.raises(_lib.ImpError):
        kernels.pool_fwd(torch.zeros(64, 256, dtype=torch.bfloat16), torch.tensor([0, 64], dtype=torch.int32), 64,
                         torch.zeros(1, 6, 256))
    net = model.IMPHotPath(n_proto=6, seed=0)
    with pytest.raises(_lib.ImpError):
        net({"img": torch.zeros(1, 8, 512), "omic": None}, lengths=[8])


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "interpretable-multimodal-prototyping_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("imp_oracle.py", ""), fn
