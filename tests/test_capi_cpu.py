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


@pytest.mark.parametrize("n_proto", [16, 32, 64])
@pytest.mark.parametrize("max_len", [4096, 16384, 120000])
def test_pool_split_plan_fills_whole_waves(max_len, n_proto):
    """The pooling kernels (pool_tc.cu) split every bag into runs of 128-row tiles; the number of runs is visible
    through the workspace query (B * nsplit partial states of PP x 258 floats, PP = 32 or 64).  On the 148-SM default
    (no GPU here) with one CTA per SM the B * nsplit CTAs must not leave a nearly empty last wave (the first version
    launched 608 CTAs on 296 slots for 32 bags: 2.05 waves), and a CTA keeps at least 4 tiles unless the bag is shorter."""
    from imp_b200 import _lib
    lib = _lib.lib()
    lib.imp_pool_fwd_workspace_bytes.restype = ctypes.c_size_t
    slots = 148
    tiles = (max_len + 127) // 128
    pp = 32 if n_proto <= 32 else 64
    for bags in (1, 2, 3, 8, 16, 32, 64, 200):
        nbytes = lib.imp_pool_fwd_workspace_bytes(bags, max_len, n_proto)
        nsplit, rem = divmod(nbytes, 4 * bags * pp * 258)
        assert rem == 0 and 1 <= nsplit <= max(1, tiles // 4), (bags, nsplit)
        ctas = bags * nsplit
        waves = -(-ctas // slots)
        tiles_per_cta = -(-tiles // nsplit)
        # no plan with fewer wave-steps exists among the admissible split counts
        best = min((-(-bags * (-(-tiles // (-(-tiles // ns))) ) // slots)) * (-(-tiles // ns) + 1)
                   for ns in range(1, max(1, min(tiles // 4, 128)) + 1))
        assert waves * (tiles_per_cta + 1) == best, (bags, nsplit, waves, tiles_per_cta, best)
        if ctas > slots:
            assert ctas / (waves * slots) > 0.6, (bags, nsplit, ctas)


def test_modularity_sweep_plan_bounds_and_waves():
    """Pair-sweep column split (modularity.cu): >= 16 and <= 256 column tiles per CTA, every column tile covered, and no
    admissible split count (up to 4x the smallest) beats the chosen one by more than 2 % in wave-steps on the 148-SM
    default.  The sharded giant bag (15 000 of 120 000 rows on one rank) must not end in a 0.4-full seventh wave."""
    from imp_b200 import _lib
    lib = _lib.lib()
    sms = 148

    def plan(own, max_len, bags):
        ns, tps = ctypes.c_int(), ctypes.c_int()
        assert lib.imp_modularity_sweep_plan(own, max_len, bags, ctypes.byref(ns), ctypes.byref(tps)) == 0
        return ns.value, tps.value

    def steps(own, max_len, bags, ns):
        col_tiles = -(-max_len // 64) + 1
        tps = -(-col_tiles // ns)
        ns_eff = -(-col_tiles // tps)
        return -(-(-(-own // 128) * bags * ns_eff) // sms) * (2 * tps + 1)

    for own, max_len, bags in [(16384, 16384, 32), (120000, 120000, 1), (15000, 120000, 1), (60000, 120000, 1),
                               (4096, 4096, 8), (700, 700, 3), (100, 100, 1), (16384, 16384, 1)]:
        ns, tps = plan(own, max_len, bags)
        col_tiles = -(-max_len // 64) + 1
        assert ns >= 1 and ns * tps >= col_tiles and (ns - 1) * tps < col_tiles, (own, max_len, bags, ns, tps)
        assert tps <= 256 or col_tiles <= 256
        assert tps >= min(16, col_tiles) - 1, (own, max_len, bags, ns, tps)
        ns_min = max(1, -(-col_tiles // 256), min(-(-2 * sms // (-(-own // 128) * bags)), max(1, col_tiles // 16)))
        mine = steps(own, max_len, bags, ns)
        for alt in range(ns_min, max(ns_min, min(4 * ns_min, max(1, col_tiles // 16))) + 1):
            assert mine * 98 <= steps(own, max_len, bags, alt) * 100 + 98, (own, max_len, bags, ns, alt)
    ns, tps = plan(15000, 120000, 1)
    ctas = -(-15000 // 128) * ns
    assert ctas / (-(-ctas // sms) * sms) > 0.95, (ns, tps, ctas)
    assert lib.imp_modularity_sweep_plan(0, 10, 1, None, None) != 0
