"""Backends the parity tests run the SAME C ABI through.

"emu"  : tests/emu/libb200conv_emu.so — engine.cu + kernels.cuh compiled by g++ against a host
         stand-in of the CUDA runtime (index arithmetic + host logic check; runs without a GPU)
"cuda" : reevr_b200/libb200conv.so — the product library, needs a B200 (pytest -m gpu)
"""
import pytest

from reevr_b200 import _lib

_cache = {}


def get_lib(name: str):
    if name not in _cache:
        if name == "emu":
            from tests.emu.build_emu import build
            _cache[name] = _lib.load(build())
        else:
            _cache[name] = _lib.default()
    return _cache[name]


BACKENDS = ["emu", pytest.param("cuda", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return get_lib(request.param)
