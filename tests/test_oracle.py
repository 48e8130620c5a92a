"""Pins the CPU oracle (oracle/partconv_oracle.c) — runs without a GPU.

1. the reference's own 58 known-answer cases (naive-convolution truth, reference tolerance);
2. the unmodified reference compiled here (oracle/_ref), same chunking, tight tolerance;
3. the committed golden fixtures (tests/golden/*.npz, generated from oracle/_ref).
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests import refcases as rc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _peak_err(a, b):
    d = np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))
    return d / max(np.max(np.abs(b.astype(np.float64))), 1e-30)


@pytest.mark.parametrize("case", rc.UNIFORM_CASES, ids=lambda c: "u-" + "-".join(map(str, c)))
def test_uniform_selftest_cases(case):
    n_in, n_ir, bmin, bmax, block = case
    x, h = rc.ramp(n_in), rc.ramp(n_ir)
    total = n_in + n_ir - 1
    chunks = rc.chunk_schedule(total, bmin, bmax, rc.GlibcRand(1))
    conv = orc.OracleUniform()
    assert conv.init(block, h)
    y = rc.drive(conv, x, total, chunks)
    truth = orc.naive_convolve(x, h)
    assert rc.reference_selftest_ok(y, truth, n_ir)
    if orc.ref_available():
        ref = orc.RefUniform()
        assert ref.init(block, h)
        yr = rc.drive(ref, x, total, chunks)
        assert _peak_err(y, yr) <= 1e-6


@pytest.mark.parametrize("case", rc.TWOSTAGE_CASES, ids=lambda c: "t-" + "-".join(map(str, c)))
def test_twostage_selftest_cases(case):
    n_in, n_ir, bmin, bmax, head, tail = case
    x, h = rc.ramp(n_in), rc.ramp(n_ir)
    total = n_in + n_ir - 1
    chunks = rc.chunk_schedule(total, bmin, bmax, rc.GlibcRand(1))
    conv = orc.OracleTwoStage()
    assert conv.init(head, tail, h)
    y = rc.drive(conv, x, total, chunks)
    truth = orc.naive_convolve(x, h)
    assert rc.reference_selftest_ok(y, truth, n_ir)
    if orc.ref_available():
        ref = orc.RefTwoStage()
        assert ref.init(head, tail, h)
        yr = rc.drive(ref, x, total, chunks)
        assert _peak_err(y, yr) <= 1e-6


def test_error_conventions():
    # FFTConvolver.cpp:97-111,157-161: zero block -> false; empty / sub-threshold IR -> true, zeros out
    c = orc.OracleUniform()
    assert not c.init(0, np.ones(4, np.float32))
    assert c.init(8, np.zeros(0, np.float32))
    assert np.all(c.process(np.ones(5, np.float32)) == 0)
    assert c.init(8, np.full(16, 5e-7, np.float32))
    assert np.all(c.process(np.ones(5, np.float32)) == 0)
    # trailing-tap trim is absolute 1e-6 and changes P (FFTConvolver.cpp:103-106)
    h = np.ones(20, np.float32)
    h[17:] = 9e-7
    assert c.init(8, h) and c.partitions == 3   # 17 taps -> ceil(17/8)
    t = orc.OracleTwoStage()
    assert not t.init(0, 8, h) and not t.init(8, 0, h)
    # non power of two rounds up (FFTConvolver.cpp:113)
    assert c.init(5, h) and c.block == 8


def test_clear_block_aligned_gives_zeros():
    # FFTConvolver.cpp:80-90 — a clear() on a block boundary silences the tail completely
    h = orc.synth_ir(1000)
    x = orc.synth_input(512)
    for conv in (orc.OracleUniform(), orc.OracleTwoStage()):
        if isinstance(conv, orc.OracleTwoStage):
            conv.init(32, 128, h)
        else:
            conv.init(64, h)
        conv.process(x)          # 512 = multiple of every block size involved
        conv.clear()
        y = conv.process(np.zeros(2048, np.float32))
        assert np.all(y == 0)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))),
                         ids=lambda p: os.path.basename(p))
def test_golden_fixture(path):
    from tests.golden.make_golden import run_case
    g = np.load(path, allow_pickle=False)
    spec = {k: (g[k].item() if g[k].ndim == 0 else g[k]) for k in g.files if k != "out"}
    y = run_case(spec, impl="oracle")
    assert y.shape == g["out"].shape
    assert _peak_err(y, g["out"]) <= 1e-6
