"""The C-ABI shared library loads and exports every symbol include/b200conv.h declares
(no compute calls — runs without a GPU)."""
import os
import re

from reevr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200conv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200conv_[a-z_0-9]+)\s*\(", src)) - {"b200conv_reduce_fn"})


def test_library_exports_every_declared_symbol():
    from reevr_b200 import build
    lib = _lib.load(build.build())          # compiles with nvcc (cross-compile, no GPU needed)
    declared = _declared_symbols()
    bound = {s[0] for s in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.b200conv_version()


def test_no_silent_fallback_without_gpu():
    # in a container without a GPU the product library must FAIL, not compute on the CPU
    import ctypes as C
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    cfg = _lib.Config(1, 0, 0, 0, 1, 0)
    h = lib.b200conv_create(C.byref(cfg))
    assert h
    ir = (C.c_float * 4)(1, 0.5, 0.25, 0.125)
    irs = (C.c_void_p * 1)(C.addressof(ir))
    lens = (C.c_size_t * 1)(4)
    assert lib.b200conv_init_uniform(h, 8, irs, lens) == -2     # B200CONV_ECUDA
    assert lib.b200conv_last_error(h) != b""
    lib.b200conv_destroy(h)
