"""TEST INFRASTRUCTURE ONLY — windowed parity checks against the CPU reference for jobs that are too long to
run through it whole (bench.py's multi-GPU `parity` objects, tests/test_refcheck.py).

The reference convolver (oracle/_ref = the unmodified FFTConvolver, else the C restatement) costs P partitions
per block, so a 28 000-block batch on a 938-partition IR (or anything on the 11 250-partition 120 s IR) cannot be
replayed in a bench run.  Two exact shortcuts:

* ``ref_window``: an output block depends on the P+1 input blocks up to it only (the IR spans P partitions), so
  a cold reference fed from block w0-P-1 reproduces blocks [w0, w0+n) of the long stream.
* ``ref_window_segmented``: by linearity y = sum_s conv(x, h_s) delayed by o_s for a split of the IR into
  partition ranges h_s = h[o_s : o_s + len_s]; each term is a SHORT reference convolver (P/S partitions) run over
  P/S + n blocks, so the cost drops from P*(P+n) to ~P*(P/S+n) block-partitions and the terms run on separate
  host threads.  Sum in float64.  This is the same sum FFTConvolver.cpp:179-187 computes, regrouped.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import oracle as orc


def _cls():
    return orc.RefUniform if orc.ref_available() else orc.OracleUniform


def kind() -> str:
    return "reference" if orc.ref_available() else "port"


def trimmed(ir: np.ndarray) -> np.ndarray:
    """FFTConvolver.cpp:103-106: trailing taps with |h| < 1e-6 do not count."""
    n = ir.size
    while n > 0 and abs(float(ir[n - 1])) < 0.000001:
        n -= 1
    return ir[:n]


def _blocks(x: np.ndarray, b0: int, b1: int, B: int) -> np.ndarray:
    """blocks [b0, b1) of the stream x, zeros for negative block indices (before the stream started)"""
    out = np.zeros((b1 - b0) * B, np.float32)
    lo = max(b0, 0)
    if b1 > lo:
        out[(lo - b0) * B:] = x[lo * B:b1 * B]
    return out


def ref_window(block: int, ir, x, w0: int, nblk: int, delay_blocks: int = 0) -> np.ndarray:
    """Reference output for blocks [w0, w0+nblk) of conv(x delayed by delay_blocks, ir), float32."""
    ir = trimmed(np.ascontiguousarray(ir, np.float32))
    if ir.size == 0:
        return np.zeros(nblk * block, np.float32)
    P = -(-ir.size // block)
    start = w0 - P - 1
    k = _cls()()
    assert k.init(block, ir)
    seg = _blocks(x, start - delay_blocks, w0 + nblk - delay_blocks, block)
    y = k.run(seg, block)
    return y[(w0 - start) * block:]


def ref_window_segmented(block: int, ir, x, w0: int, nblk: int, nseg: int = 32, threads: int = 16) -> np.ndarray:
    ir = trimmed(np.ascontiguousarray(ir, np.float32))
    P = -(-ir.size // block)
    nseg = max(1, min(nseg, P))
    per = -(-P // nseg)
    jobs = [(p0, min(P, p0 + per)) for p0 in range(0, P, per)]

    def one(job):
        p0, p1 = job
        return ref_window(block, ir[p0 * block:p1 * block], x, w0, nblk, delay_blocks=p0).astype(np.float64)

    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        parts = list(ex.map(one, jobs))
    return np.sum(parts, axis=0)


def peak_err(y, ref) -> float:
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(np.asarray(y, np.float64) - ref)) / max(float(np.max(np.abs(ref))), 1e-30))
