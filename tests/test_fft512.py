"""Register-resident B = 512 FFT kernels (kernels_fft512.cuh): selected for launches of >= 32 transforms; checked
against the oracle through every epilogue form (aligned whole blocks, partial blocks, look-ahead ring destination,
added tail rings) and against the shared-memory Stockham kernels they replace."""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import Engine
from tests.backends import lib  # noqa: F401

TOL = 1e-5


def peak_err(y, ref):
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def test_uniform_512_batches(lib):
    irs = [orc.synth_ir(512 * 9 - 100, c) for c in range(2)]
    n = 512 * 70
    x = [orc.synth_input(n, c) for c in range(2)]
    e = Engine(2, lib=lib)
    assert e.init_uniform(512, irs)
    # whole blocks (fast epilogue), then ragged calls: open block + many blocks, ending mid-block (generic epilogue)
    chunks = [512 * 40, 300, 512 * 20 + 77, 512 * 9 + 135]
    assert sum(chunks) == n
    outs = [[], []]
    pos = 0
    for k in chunks:
        ys = e.process([a[pos:pos + k] for a in x])
        for c in range(2):
            outs[c].append(ys[c])
        pos += k
    for c in range(2):
        o = orc.OracleUniform()
        o.init(512, irs[c])
        assert peak_err(np.concatenate(outs[c]), o.process(x[c])) <= TOL


def test_impulse_through_every_bin_and_partition(lib):
    """unit impulses at different offsets inside a block reproduce the IR: exercises every twiddle / index of both
    transforms (a wrong lane mapping cannot cancel out)"""
    ir = orc.synth_ir(512 * 3)
    e = Engine(1, lib=lib)
    assert e.init_uniform(512, [ir])
    n = 512 * 40
    for off in (0, 1, 255, 256, 511):
        e.clear()
        x = np.zeros(n, np.float32)
        x[off] = 1.0
        y = e.process([x])[0]
        want = np.zeros(n, np.float32)
        want[off:off + ir.size] = ir
        assert np.max(np.abs(y - want)) <= 2e-6, off


def test_two_stage_with_512_blocks(lib):
    # head 512 with an added tail ring (generic epilogue with n_add > 0) ...
    ir = orc.synth_ir(4096 * 2 + 4096 * 3 + 55)
    x = orc.synth_input(512 * 64)
    e = Engine(1, lib=lib)
    assert e.init_twostage(512, 4096, [ir])
    o = orc.OracleTwoStage()
    o.init(512, 4096, ir)
    assert peak_err(e.process([x])[0], o.process(x)) <= TOL
    # ... and a 512-block TAIL stage writing into its look-ahead ring (masked destination), 40 tail blocks per call
    ir2 = orc.synth_ir(1024 + 512 * 6)
    x2 = orc.synth_input(512 * 80)
    e2 = Engine(1, lib=lib)
    assert e2.init_twostage(64, 512, [ir2])
    o2 = orc.OracleTwoStage()
    o2.init(64, 512, ir2)
    y2 = np.concatenate([e2.process([x2[i:i + 512 * 40]])[0] for i in range(0, x2.size, 512 * 40)])
    assert peak_err(y2, o2.process(x2)) <= TOL
