"""ctypes binding of libimp_sm100.so (include/imp_hotpath.h).  There is no CPU or eager
fallback: if the library is missing it is built with nvcc, and if that fails import fails."""
from __future__ import annotations

import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "imp_hotpath.h")
LIB_PATH = os.path.join(HERE, "libimp_sm100.so")

_lib = None


class ImpError(RuntimeError):
    pass


def header_symbols(header: str = HEADER):
    """Names of every function include/imp_hotpath.h declares."""
    text = open(header).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(imp_[a-z0-9_]+)\s*\(", text)))


def header_abi_version(header: str = HEADER) -> int:
    m = re.search(r"#define\s+IMP_ABI_VERSION\s+(\d+)", open(header).read())
    if not m:
        raise ImpError("IMP_ABI_VERSION not found in %s" % header)
    return int(m.group(1))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        from . import build as _build
        try:                                   # digest-based and incremental: a stale library is rebuilt, a fresh one is kept
            _build.build()
        except RuntimeError:
            if not os.path.exists(LIB_PATH):   # no nvcc and no prebuilt library: there is nothing to fall back to
                raise
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.imp_last_error.restype = ctypes.c_char_p
        _lib.imp_abi_version.restype = ctypes.c_int
        want = header_abi_version()
        got = int(_lib.imp_abi_version())
        if got != want:
            raise ImpError("libimp_sm100.so reports ABI %d, include/imp_hotpath.h declares %d: rebuild the library" % (got, want))
        for name in header_symbols():
            fn = getattr(_lib, name)          # AttributeError if the .so lacks a declared symbol
            if name.endswith(("_bytes", "_floats")):
                fn.restype = ctypes.c_size_t
    return _lib


def _conv(a):
    import torch
    if a is None:
        return ctypes.c_void_p(0)
    if isinstance(a, torch.Tensor):
        return ctypes.c_void_p(a.data_ptr())
    if isinstance(a, bool):
        return ctypes.c_int(int(a))
    if isinstance(a, int):
        return ctypes.c_int(a)
    if isinstance(a, float):
        return ctypes.c_float(a)
    return a


def call(name: str, *args) -> None:
    """Call an int-returning entry point; raise ImpError with the library's message on failure."""
    rc = getattr(lib(), name)(*[_conv(a) for a in args])
    if rc != 0:
        raise ImpError("%s failed (%d): %s" % (name, rc, lib().imp_last_error().decode()))


def query(name: str, *args) -> int:
    return int(getattr(lib(), name)(*[_conv(a) for a in args]))


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count() -> int:
    fn = lib().imp_launch_count
    fn.restype = ctypes.c_longlong
    return int(fn())


def profile_enable(on: bool) -> None:
    lib().imp_profile_enable(ctypes.c_int(1 if on else 0))


def profile_collect(max_records: int = 65536):
    """-> list of (kernel name, ms) for the launches recorded since the last collect."""
    names = (ctypes.c_char_p * max_records)()
    ms = (ctypes.c_float * max_records)()
    n = lib().imp_profile_collect(names, ms, ctypes.c_int(max_records))
    if n < 0:
        raise ImpError("imp_profile_collect: CUDA error")
    return [(names[i].decode(), float(ms[i])) for i in range(n)]
