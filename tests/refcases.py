"""The (shape, chunking) cases of the reference's own self-test, restated as data.

Source: libs/FFTConvolver/test/Test.cpp:256-288 (uniform, 29 cases) and :297-329 (two-stage,
29 cases).  Inputs are the ramps in[i] = ir[i] = 0.1*(i+1) (:77-87); process() is called with
random chunk lengths in [bmin, bmax] drawn from glibc rand() (:106), zero input after the
signal ends, until in+ir-1 samples are produced (:104-124).
"""
import ctypes

import numpy as np

# (inputSize, irSize, blockSizeMin, blockSizeMax, blockSizeConvolver)
UNIFORM_CASES = [
    (1, 1, 1, 1, 1), (2, 2, 2, 2, 2), (3, 3, 3, 3, 3),
    (3, 2, 2, 2, 2), (4, 2, 2, 2, 2), (4, 3, 2, 2, 2), (9, 4, 3, 3, 2), (171, 7, 5, 5, 5),
    (1979, 17, 7, 7, 5), (100, 10, 3, 5, 5), (123, 45, 12, 34, 34),
    (2, 3, 2, 2, 2), (2, 4, 2, 2, 2), (3, 4, 2, 2, 2), (4, 9, 3, 3, 3), (7, 171, 5, 5, 5),
    (17, 1979, 7, 7, 7), (10, 100, 3, 5, 5), (45, 123, 12, 34, 34),
    (100000, 1234, 100, 128, 128), (100000, 1234, 100, 256, 256), (100000, 1234, 100, 512, 512),
    (100000, 1234, 100, 1024, 1024), (100000, 1234, 100, 2048, 2048),
    (100000, 4321, 100, 128, 128), (100000, 4321, 100, 256, 256), (100000, 4321, 100, 512, 512),
    (100000, 4321, 100, 1024, 1024), (100000, 4321, 100, 2048, 2048),
]

# (inputSize, irSize, blockSizeMin, blockSizeMax, blockSizeHead, blockSizeTail)
TWOSTAGE_CASES = [
    (1, 1, 1, 1, 1, 1), (2, 2, 2, 2, 2, 2), (3, 3, 3, 3, 3, 3),
    (3, 2, 2, 2, 2, 4), (4, 2, 2, 2, 2, 4), (4, 3, 2, 2, 2, 4), (9, 4, 3, 3, 2, 4),
    (171, 7, 5, 5, 5, 10), (1979, 17, 7, 7, 5, 10), (100, 10, 3, 5, 5, 10), (123, 45, 12, 34, 34, 68),
    (2, 3, 2, 2, 1, 2), (2, 4, 2, 2, 1, 2), (3, 4, 2, 2, 1, 2), (4, 9, 3, 3, 2, 4),
    (7, 171, 5, 5, 2, 16), (17, 1979, 7, 7, 4, 16), (10, 100, 3, 5, 1, 4), (45, 123, 12, 34, 4, 32),
    (100000, 1234, 100, 128, 128, 4096), (100000, 1234, 100, 256, 256, 4096),
    (100000, 1234, 100, 512, 512, 4096), (100000, 1234, 100, 1024, 1024, 4096),
    (100000, 1234, 100, 2048, 2048, 4096),
    (100000, 4321, 100, 128, 128, 4096), (100000, 4321, 100, 256, 256, 4096),
    (100000, 4321, 100, 512, 512, 4096), (100000, 4321, 100, 1024, 1024, 4096),
    (100000, 4321, 100, 2048, 2048, 4096),
]


def ramp(n: int) -> np.ndarray:
    """0.1f * float(i+1) evaluated in float32 as the reference does (Test.cpp:79,86)."""
    return (np.float32(0.1) * np.arange(1, n + 1, dtype=np.float32)).astype(np.float32)


class GlibcRand:
    """glibc rand() stream; the reference never seeds it, i.e. srand(1)."""

    def __init__(self, seed: int = 1):
        self._libc = ctypes.CDLL(None)
        self._libc.rand.restype = ctypes.c_int
        self._libc.srand(seed)

    def __call__(self) -> int:
        return int(self._libc.rand())


def chunk_schedule(total_out: int, bmin: int, bmax: int, rnd) -> list:
    """Chunk lengths the reference test would use to produce `total_out` samples (Test.cpp:104-124)."""
    out, done = [], 0
    while done < total_out:
        b = bmin + (rnd() % (1 + (bmax - bmin)))
        n = min(total_out - done, b)
        out.append(n)
        done += n
    return out


def drive(conv, x: np.ndarray, total_out: int, chunks: list) -> np.ndarray:
    """Feed x (then zeros) through conv.process() in the given chunks; returns total_out samples."""
    xin = np.zeros(total_out, dtype=np.float32)
    xin[: x.size] = x
    y = np.empty(total_out, dtype=np.float32)
    pos = 0
    for n in chunks:
        y[pos:pos + n] = conv.process(xin[pos:pos + n])
        pos += n
    assert pos == total_out
    return y


def reference_selftest_ok(out: np.ndarray, truth: np.ndarray, ir_len: int) -> bool:
    """Pass criterion of the reference self-test (Test.cpp:127-148)."""
    a = out.astype(np.float64)
    b = truth.astype(np.float64)
    abs_tol = 0.001 * ir_len
    rel_tol = 0.0001 * np.log(float(ir_len)) if ir_len > 1 else 0.0
    m = (np.abs(a) > 1.0) & (np.abs(b) > 1.0)
    abs_err = np.abs(a - b)[m]
    rel_err = abs_err / b[m]
    return not np.any((rel_err > rel_tol) & (abs_err > abs_tol))
