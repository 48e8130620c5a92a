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
    if _needs_build(_ORACLE_SO, os.path.join(_HERE, "partconv_oracle.c")) or \
            _needs_build(_ORACLE_SO, os.path.join(_HERE, "chain_oracle.c")):
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
    # send / wet chain (chain_oracle.c)
    lib.oc_filter_create.restype = C.c_void_p
    lib.oc_filter_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    lib.oc_filter_destroy.argtypes = [C.c_void_p]
    lib.oc_filter_run.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t]
    lib.oc_filter_coeff.restype = C.c_float
    lib.oc_filter_coeff.argtypes = [C.c_float, C.c_float]
    lib.oc_chain_create.restype = C.c_void_p
    lib.oc_chain_create.argtypes = [C.c_float, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float]
    lib.oc_chain_destroy.argtypes = [C.c_void_p]
    lib.oc_chain_send.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_size_t]
    lib.oc_chain_wet.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, _f32p, _f32p, _f32p, C.c_size_t]
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


_REF_FILTER_SO = os.path.join(_HERE, "_ref", "libreffilter.so")


def ref_filter_available() -> bool:
    if os.path.exists(_REF_FILTER_SO):
        return True
    if os.path.isdir("/root/reference/src/dsp"):
        build()
        return os.path.exists(_REF_FILTER_SO)
    return False


class OracleFilter:
    """C restatement of REEV-R's Filter (src/dsp/Filter.cpp): slope 0/1/2 = 6/12/24 dB, mode 0/1/2 = LP/BP/HP."""

    def __init__(self, slope: int, mode: int, srate: float, freq: float, q: float):
        self._l = _lib("oc")
        self._h = self._l.oc_filter_create(slope, mode, srate, freq, q)

    def run(self, x) -> np.ndarray:
        x = _as_f32(x)
        y = np.empty_like(x)
        self._l.oc_filter_run(self._h, x, y, x.size)
        return y

    def __del__(self):
        try:
            self._l.oc_filter_destroy(self._h)
        except Exception:
            pass


class RefFilter:
    """The unmodified reference Filter (oracle/_ref/libreffilter.so)."""

    def __init__(self, slope: int, mode: int, srate: float, freq: float, q: float):
        if "reffilter" not in _libs:
            if not ref_filter_available():
                raise RuntimeError("oracle/_ref/libreffilter.so not built and /root/reference absent")
            l = C.CDLL(_REF_FILTER_SO)
            l.ref_filter_create.restype = C.c_void_p
            l.ref_filter_create.argtypes = [C.c_int, C.c_int]
            l.ref_filter_destroy.argtypes = [C.c_void_p]
            l.ref_filter_init.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
            l.ref_filter_reset.argtypes = [C.c_void_p, C.c_float]
            l.ref_filter_run.argtypes = [C.c_void_p, _f32p, _f32p, C.c_size_t]
            l.ref_filter_coeff.restype = C.c_float
            l.ref_filter_coeff.argtypes = [C.c_float, C.c_float]
            _libs["reffilter"] = l
        self._l = _libs["reffilter"]
        self._h = self._l.ref_filter_create(slope, mode)
        self._l.ref_filter_init(self._h, srate, freq, q)
        self._l.ref_filter_reset(self._h, 0.0)

    def run(self, x) -> np.ndarray:
        x = _as_f32(x)
        y = np.empty_like(x)
        self._l.ref_filter_run(self._h, x, y, x.size)
        return y

    def __del__(self):
        try:
            self._l.ref_filter_destroy(self._h)
        except Exception:
            pass


def filter_coeff(freq: float, srate: float, ref: bool = False) -> float:
    if ref:
        RefFilter(0, 0, srate, freq, 0.5)          # loads the library
        return float(_libs["reffilter"].ref_filter_coeff(freq, srate))
    return float(_lib("oc").oc_filter_coeff(freq, srate))


class OracleChain:
    """C restatement of the send / wet chain of processBlock (src/PluginProcessor.cpp:1639-1653, 1766-1790, 1832-1876)."""

    def __init__(self, srate, lowcut_hz, lowcut_slope, highcut_hz, highcut_slope, predelay, width, drygain, wetgain,
                 delay_size: int = 0):
        self._l = _lib("oc")
        self._h = self._l.oc_chain_create(srate, lowcut_hz, lowcut_slope, highcut_hz, highcut_slope, predelay,
                                          delay_size or max(2 * predelay, 1), width, drygain, wetgain)

    def send(self, dryL, dryR, ysend):
        dryL, dryR, ysend = _as_f32(dryL), _as_f32(dryR), _as_f32(ysend)
        a, b = np.empty_like(dryL), np.empty_like(dryR)
        self._l.oc_chain_send(self._h, dryL, dryR, ysend, a, b, dryL.size)
        return a, b

    def wet(self, dryL, dryR, LL, RR, LR, RL, yrev):
        arrs = [_as_f32(v) for v in (dryL, dryR, LL, RR)]
        lr = _as_f32(LR) if LR is not None else None
        rl = _as_f32(RL) if RL is not None else None
        yrev = _as_f32(yrev)
        oL, oR = np.empty_like(arrs[0]), np.empty_like(arrs[1])
        self._l.oc_chain_wet(self._h, *arrs, lr.ctypes.data if lr is not None else None,
                             rl.ctypes.data if rl is not None else None, yrev, oL, oR, arrs[0].size)
        return oL, oR

    def __del__(self):
        try:
            self._l.oc_chain_destroy(self._h)
        except Exception:
            pass


def apply_decay(ir, lut, srate: float) -> np.ndarray:
    """C restatement of Impulse::applyDecay (src/dsp/Impulse.cpp:602-648); lut has 2049 entries."""
    buf = np.array(ir, dtype=np.float32, copy=True)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    assert lut.size == 2049
    if buf.size:
        _lib("oc").oc_apply_decay(buf, buf.size, lut, float(srate))
    return buf


def ir_shape(irs, autogain=True, reverse=False, trim_left=0.0, trim_right=0.0, gain=1.0, lut=None, srate=48000.0,
             clip=True, attack=0.0, decay=0.0):
    """C restatement of the device-resident subset of Impulse::recalcImpulse (chain_oracle.c::oc_ir_shape); returns the
    shaped channels (possibly shorter: trim)."""
    lib = _lib("oc")
    lib.oc_ir_shape.restype = C.c_size_t
    lib.oc_ir_shape.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                C.c_double, C.c_int, C.c_float, C.c_float]
    bufs = [np.array(a, dtype=np.float32, copy=True) for a in irs]
    n = bufs[0].size
    ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    l = None if lut is None else np.ascontiguousarray(lut, dtype=np.float64)
    m = lib.oc_ir_shape(ptrs, len(bufs), n, int(autogain), int(reverse), trim_left, trim_right, gain,
                        l.ctypes.data if l is not None else None, float(srate), int(clip), attack, decay)
    return [b[:m].copy() for b in bufs]


def ref_apply_decay(ir, lut, srate: float) -> np.ndarray:
    """The same STFT loop driven through the reference's own compiled AudioFFT (oracle/ref_shim.cpp::ref_stft_decay)."""
    lib = _lib("ref")
    lib.ref_stft_decay.restype = None
    lib.ref_stft_decay.argtypes = [_f32p, C.c_size_t, np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS"), C.c_double, _f32p]
    buf = np.array(ir, dtype=np.float32, copy=True)
    lut = np.ascontiguousarray(lut, dtype=np.float64)
    if buf.size:
        lib.ref_stft_decay(buf, buf.size, lut, float(srate), decay_window())
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
