"""The one-launch real-time path (k_rt_block: cluster kernel, zero-copy I/O, tail blocks on the low-priority
stream) against the oracle and against the multi-kernel path it replaces, for the shapes a REEV-R audio callback
produces (StereoConvolver.cpp:8-42): host blocks that are / are not a power of two, stereo and quad, mixdown."""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import Engine, StereoConvolver
from tests.backends import lib  # noqa: F401

TOL = 1e-5


def peak_err(y, ref):
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def stream(e, xs, chunks):
    outs = [[] for _ in range(len(e.process([a[:0] for a in xs])) or len(xs))]
    pos = 0
    for k in chunks:
        ys = e.process([a[pos:pos + k] for a in xs])
        if len(outs) != len(ys):
            outs = [[] for _ in ys]
        for c, y in enumerate(ys):
            outs[c].append(y)
        pos += k
    return [np.concatenate(o) for o in outs]


@pytest.mark.parametrize("head,tail,host_block", [(64, 512, 64), (128, 1024, 100), (512, 2048, 512), (256, 512, 37)])
def test_two_stage_callback_sequence(lib, head, tail, host_block):
    irs = [orc.synth_ir(2 * tail + 5 * tail + 123, c) for c in range(2)]
    n = host_block * 150
    xs = [orc.synth_input(n, c) for c in range(2)]
    chunks = [host_block] * 150
    res = {}
    for rt in (1, 0):
        e = Engine(2, lib=lib)
        e.set_option("rt", rt)
        assert e.init_twostage(head, tail, irs)
        l0 = e.launch_count
        res[rt] = stream(e, xs, chunks)
        res[("launches", rt)] = e.launch_count - l0
    for c in range(2):
        o = orc.OracleTwoStage()
        o.init(head, tail, irs[c])
        ref = o.process(xs[c])
        assert peak_err(res[1][c], ref) <= TOL
        assert peak_err(res[0][c], ref) <= TOL
    assert res[("launches", 1)] < res[("launches", 0)]
    if host_block == head:          # every call stays inside the open block: one launch per call + the tail blocks
        tail_blocks = n // tail
        assert res[("launches", 1)] <= 150 + 4 * tail_blocks + 2


def test_uniform_long_ir_uses_a_multi_cta_cluster(lib):
    # 600 partitions of 256: 2.4 MB of spectra per convolver -> several CTAs per convolver, DSMEM tile gather
    irs = [orc.synth_ir(256 * 600 - 9, c) for c in range(2)]
    xs = [orc.synth_input(256 * 40, c) for c in range(2)]
    e = Engine(2, lib=lib)
    assert e.init_uniform(256, irs)
    ys = stream(e, xs, [256] * 20 + [100, 156] + [256] * 19)
    for c in range(2):
        o = orc.OracleUniform()
        o.init(256, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL


def test_uniform_very_long_ir_uses_the_split_mode(lib):
    """head stage beyond one cluster's reach (1100 partitions of 256 = 4.5 MB per convolver): front kernel, all-SM TMA
    sweep, back kernel — same zero-copy I/O, incl. a ragged pair of calls and the device mixdown"""
    irs = [orc.synth_ir(256 * 1100 - 9, c) for c in range(2)]
    xs = [orc.synth_input(256 * 30, c) for c in range(2)]
    e = Engine(2, lib=lib)
    assert e.init_uniform(256, irs)
    l0 = e.launch_count
    ys = stream(e, xs, [256] * 12 + [100, 156] + [256] * 17)
    assert e.launch_count - l0 <= 31 * 3 + 2           # front + sweep + back per call, nothing else
    for c in range(2):
        o = orc.OracleUniform()
        o.init(256, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL
    e.set_routing([0, 1], [[1, 1], [1, -1]])
    m = stream(e, xs, [256] * 4)
    e2 = Engine(2, lib=lib)
    e2.set_option("rt", 0)
    assert e2.init_uniform(256, irs)
    stream(e2, xs, [256 * 30])
    r = stream(e2, xs, [256] * 4)
    assert peak_err(m[0], r[0] + r[1]) <= TOL and peak_err(m[1], r[0] - r[1]) <= TOL


def test_quad_with_device_mixdown(lib):
    sc = StereoConvolver(lib=lib)
    sc.prepare(128)
    irs = [orc.synth_ir(30000, c) for c in range(4)]          # LL, RR, LR, RL
    sc.loadImpulse(*irs)
    sc.enable_device_mixdown(true_stereo=True)
    n = 128 * 90
    L, R = orc.synth_input(n, 0), orc.synth_input(n, 1)
    wl, wr = np.empty_like(L), np.empty_like(R)
    for i in range(90):
        seg = slice(128 * i, 128 * (i + 1))
        wl[seg], wr[seg] = sc.process_mixed(L[seg], R[seg])
    outs = []
    for ir, src in zip(irs, (L, R, L, R)):
        o = orc.OracleTwoStage()
        o.init(128, 8192, ir)
        outs.append(o.process(src))
    LL, RR, LR, RL = outs
    assert peak_err(wl, LL + RL) <= TOL and peak_err(wr, RR + LR) <= TOL


def test_realtime_and_batch_calls_interleave(lib):
    """batch calls (multi-kernel path) in between real-time calls, clear() in the middle, a tail block in flight"""
    ir = orc.synth_ir(64 * 2 * 4 + 256 * 7)
    x = orc.synth_input(64 * 400)
    e = Engine(1, lib=lib)
    assert e.init_twostage(64, 256, [ir])
    o = orc.OracleTwoStage()
    o.init(64, 256, ir)
    chunks = [64] * 9 + [64 * 30] + [64] * 3 + [10, 54] + [64 * 5 + 7] + [57] + [64] * 20
    y = stream(e, [x], chunks)[0]
    n = sum(chunks)
    assert peak_err(y, o.process(x[:n])) <= TOL
    e.clear()
    o.clear()
    y2 = stream(e, [x[n:]], [64] * 40)[0]
    assert peak_err(y2, o.process(x[n:n + 64 * 40])) <= TOL


def test_eight_convolvers_in_one_cluster(lib):
    """config 4's channel count through the real-time path: 8 convolvers x 2 CTAs = one 16-CTA (non-portable) cluster"""
    irs = [orc.synth_ir(128 * 40 - 3, c) for c in range(8)]
    xs = [orc.synth_input(128 * 24, c) for c in range(8)]
    e = Engine(8, lib=lib)
    assert e.init_uniform(128, irs)
    l0 = e.launch_count
    ys = stream(e, xs, [128] * 24)
    assert e.launch_count - l0 == 24                     # one launch per call
    for c in range(8):
        o = orc.OracleUniform()
        o.init(128, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL
