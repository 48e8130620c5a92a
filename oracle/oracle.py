"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the CPU oracles.

* ``OracleUniform`` / ``OracleTwoStage``: the plain-C restatement (oracle/partconv_oracle.c).
* ``RefUniform`` / ``RefTwoStage``: the UNMODIFIED reference classes compiled from
  /root/reference into oracle/_ref/libreforacle.so (oracle/ref_shim.cpp, oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (reevr_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libreforacle.so")

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference is present, _ref/libreforacle.so."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _needs_build(so: str, src: str) -> bool:
    return (not os.path.exists(so)) or os.path.getmtime(so) < os.path.getmtime(src)


def _load_oracle() -> C.CDLL:
    if _needs_build(_ORACLE_SO, os.path.join(_HERE, "partconv_oracle.c")):
        build()
    lib = C.CDLL(_ORACLE_SO)
    for kind in ("uniform", "twostage"):
        getattr(lib, f"oc_{kind}_create").restype = C.c_void_p
        getattr(lib, f"oc_{kind}_create").argtypes = []
        for fn in ("destroy", "clear", "reset"):
            f = getattr(lib, f"oc_{kind}_{fn}")
            f.restype = None
            f.argtypes = [C.c_void_p]
        f = getattr(lib, f"oc_{kind}_process")
        f.restype = None
        f.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t]
        f = getattr(lib, f"oc_{kind}_run")
        f.restype = None
        f.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t, C.c_size_t]
    lib.oc_uniform_init.restype = C.c_int
    lib.oc_uniform_init.argtypes = [C.c_void_p, C.c_size_t, _f32p, C.c_size_t]
    lib.oc_twostage_init.restype = C.c_int
    lib.oc_twostage_init.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t]
    lib.oc_naive_convolve.restype = None
    lib.oc_naive_convolve.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p]
    lib.oc_apply_decay.restype = None
    lib.oc_apply_decay.argtypes = [_f32p, C.c_size_t, np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"), C.c_double]
    lib.oc_decay_window.restype = None
    lib.oc_decay_window.argtypes = [_f32p]
    lib.oc_uniform_partitions.restype = C.c_size_t
    lib.oc_uniform_partitions.argtypes = [C.c_void_p]
    lib.oc_uniform_block.restype = C.c_size_t
    lib.oc_uniform_block.argtypes = [C.c_void_p]
    return lib


def ref_available() -> bool:
    if os.path.exists(_REF_SO):
        return True
    if os.path.isdir("/root/reference/libs/FFTConvolver"):
        build()
        return os.path.exists(_REF_SO)
    return False


def _load_ref() -> C.CDLL:
    if not ref_available():
        raise RuntimeError("oracle/_ref/libreforacle.so not built and /root/reference absent")
    lib = C.CDLL(_REF_SO)
    for kind in ("uniform", "twostage"):
        getattr(lib, f"ref_{kind}_create").restype = C.c_void_p
        getattr(lib, f"ref_{kind}_create").argtypes = []
        for fn in ("destroy", "clear", "reset"):
            f = getattr(lib, f"ref_{kind}_{fn}")
            f.restype = None
            f.argtypes = [C.c_void_p]
        f = getattr(lib, f"ref_{kind}_process")
        f.restype = None
        f.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t]
        f = getattr(lib, f"ref_{kind}_run")
        f.restype = None
        f.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t, C.c_size_t]
    lib.ref_uniform_init.restype = C.c_int
    lib.ref_uniform_init.argtypes = [C.c_void_p, C.c_size_t, _f32p, C.c_size_t]
    lib.ref_twostage_init.restype = C.c_int
    lib.ref_twostage_init.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t]
    return lib


_libs: dict = {}


def _lib(which: str) -> C.CDLL:
    if which not in _libs:
        _libs[which] = _load_oracle() if which == "oc" else _load_ref()
    return _libs[which]


def _as_f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.size == 0:  # ctypes needs a valid pointer even for empty arrays
        a = np.zeros(1, dtype=np.float32)[:0]
    return a


class _Base:
    _which = "oc"
    _kind = "uniform"

    def __init__(self):
        self._l = _lib(self._which)
        self._h = self._fn("create")()

    def _fn(self, name):
        prefix = "oc" if self._which == "oc" else "ref"
        return getattr(self._l, f"{prefix}_{self._kind}_{name}")

    def process(self, x) -> np.ndarray:
        x = _as_f32(x)
        y = np.empty(max(x.size, 1), dtype=np.float32)[: x.size]
        if x.size:
            self._fn("process")(self._h, x, y, x.size)
        return y

    def run(self, x, chunk: int) -> np.ndarray:
        """process() in `chunk`-sized calls, loop on the C side (x.size % chunk must be 0)."""
        x = _as_f32(x)
        assert x.size % chunk == 0
        y = np.empty_like(x)
        self._fn("run")(self._h, x, y, chunk, x.size // chunk)
        return y

    def clear(self):
        self._fn("clear")(self._h)

    def reset(self):
        self._fn("reset")(self._h)

    def __del__(self):
        try:
            self._fn("destroy")(self._h)
        except Exception:
            pass


class OracleUniform(_Base):
    """C restatement of fftconvolver::FFTConvolver (FFTConvolver.h:62-80)."""

    def init(self, block: int, ir) -> bool:
        ir = _as_f32(ir)
        pad = ir if ir.size else np.zeros(1, np.float32)
        return bool(self._fn("init")(self._h, block, pad, ir.size))

    @property
    def partitions(self) -> int:
        return int(self._l.oc_uniform_partitions(self._h))

    @property
    def block(self) -> int:
        return int(self._l.oc_uniform_block(self._h))


class OracleTwoStage(_Base):
    """C restatement of fftconvolver::TwoStageFFTConvolver (TwoStageFFTConvolver.h:65-83)."""
    _kind = "twostage"

    def init(self, head: int, tail: int, ir) -> bool:
        ir = _as_f32(ir)
        pad = ir if ir.size else np.zeros(1, np.float32)
        return bool(self._fn("init")(self._h, head, tail, pad, ir.size))


class RefUniform(OracleUniform):
    """The unmodified reference FFTConvolver (oracle/_ref)."""
    _which = "ref"

    @property
    def partitions(self):  # private state of the reference class, not exposed
        raise AttributeError("partitions")

    @property
    def block(self):
        raise AttributeError("block")


class RefTwoStage(OracleTwoStage):
    """The unmodified reference TwoStageFFTConvolver (oracle/_ref)."""
    _which = "ref"


def apply_decay(ir, lut, srate: float) -> np.ndarray:
    """C restatement of Impulse::applyDecay (src/dsp/Impulse.cpp:602-648); lut has 2049 entries."""
    buf = np.array(ir, dtype=np.float32, copy=True)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    assert lut.size == 2049
    if buf.size:
        _lib("oc").oc_apply_decay(buf, buf.size, lut, float(srate))
    return buf


def decay_window() -> np.ndarray:
    w = np.empty(4096, np.float32)
    _lib("oc").oc_decay_window(w)
    return w


def naive_convolve(x, h) -> np.ndarray:
    """float32 direct convolution, the reference self-test's truth (test/Test.cpp:32-66)."""
    x = _as_f32(x)
    h = _as_f32(h)
    out = np.zeros(x.size + h.size - 1, dtype=np.float32)
    _lib("oc").oc_naive_convolve(x, x.size, h, h.size, out)
    return out


# Synthetic signals of SURVEY.md §8(d) live in the (dependency-free) product module so that
# bench.py's GPU arm does not have to import anything from oracle/.
import sys as _sys

_sys.path.insert(0, os.path.dirname(_HERE))
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402,F401
