"""Time-slice sharding (b200conv_process_sliced / _device_sliced): G full convolvers, each producing one contiguous
time slice of a block-aligned call with no exchange; afterwards every handle is in the state the whole call leaves.
Checked against the oracle over SEQUENCES of calls (the history crosses call boundaries), mixed with plain
process() calls, for T < P, G > T, several channels; emulation here, the same tests on the GPU (-m gpu)."""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import B200ConvError, Engine
from tests.backends import lib  # noqa: F401

TOL = 1e-5


def peak_err(y, ref):
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def run_sliced_sequence(lib, G, B, irs, xs, calls, max_batch_blocks=0):
    """calls: list of (n_samples, sliced?)  -> per-channel output of the whole stream"""
    C = len(irs)
    engs = [Engine(C, max_batch_blocks=max_batch_blocks, lib=lib) for _ in range(G)]
    for e in engs:
        assert e.init_uniform(B, irs)
    outs = [np.zeros_like(x) for x in xs]
    pos = 0
    for n, sliced in calls:
        seg_in = [np.ascontiguousarray(x[pos:pos + n]) for x in xs]
        if sliced:
            seg_out = [np.full(n, np.nan, np.float32) for _ in range(C)]
            for g in range(G):
                engs[g].process_sliced(seg_in, seg_out, g, G)
            for c in range(C):
                assert not np.isnan(seg_out[c]).any(), "a slice was not written"
                outs[c][pos:pos + n] = seg_out[c]
        else:
            ys = [e.process(seg_in) for e in engs]            # every replica advances; all must agree
            for g in range(1, G):
                for c in range(C):
                    assert peak_err(ys[g][c], ys[0][c]) <= 1e-6
            for c in range(C):
                outs[c][pos:pos + n] = ys[0][c]
        pos += n
    for e in engs:
        e.close()
    return outs


@pytest.mark.parametrize("G", [1, 2, 3, 8])
def test_sliced_calls_continue_one_stream(lib, G):
    B, P = 64, 23
    irs = [orc.synth_ir(P * B - 7, c) for c in range(2)]
    T = [100, 40, 9, 64, 3]                                   # blocks per call: T >> P, T ~ P, T < P, T < G
    calls = [(t * B, True) for t in T]
    n = sum(c[0] for c in calls)
    xs = [orc.synth_input(n, c) for c in range(2)]
    ys = run_sliced_sequence(lib, G, B, irs, xs, calls, max_batch_blocks=24)
    for c in range(2):
        o = orc.OracleUniform()
        o.init(B, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL


@pytest.mark.parametrize("G", [2, 4])
def test_sliced_and_plain_calls_mix(lib, G):
    B = 32
    ir = orc.synth_ir(50 * B)
    calls = [(64 * B, True), (777, False), (B - 777 % B, False), (10 * B, True), (5 * B + 3, False)]
    n = sum(c[0] for c in calls)
    x = orc.synth_input(n)
    y = run_sliced_sequence(lib, G, B, [ir], [x], calls)[0]
    o = orc.OracleUniform()
    o.init(B, ir)
    assert peak_err(y, o.process(x)) <= TOL


def test_sliced_without_tail_on_the_later_ranks(lib):
    """steady batch job: ranks whose slice starts >= P blocks into the call skip the tail (slice_keep_tail = 0)"""
    B, P, G = 64, 9, 4
    irs = [orc.synth_ir(P * B - 3)]
    T = 60
    xs = [orc.synth_input(3 * T * B)]
    engs = [Engine(1, lib=lib) for _ in range(G)]
    for g, e in enumerate(engs):
        assert e.init_uniform(B, irs)
        if g > 0:
            e.set_option("slice_keep_tail", 0)
    out = np.zeros_like(xs[0])
    for call in range(3):
        seg = [np.ascontiguousarray(xs[0][call * T * B:(call + 1) * T * B])]
        o = [np.full(T * B, np.nan, np.float32)]
        for g in range(G):
            engs[g].process_sliced(seg, o, g, G)
        out[call * T * B:(call + 1) * T * B] = o[0]
    ref = orc.OracleUniform()
    ref.init(B, irs[0])
    assert peak_err(out, ref.process(xs[0])) <= TOL


def test_sliced_device_resident_matches_host_path(lib):
    import ctypes as C
    if b"EMULATED" in lib.b200conv_version():
        def dev(a):
            return a, a.ctypes.data

        def back(a):
            return a
    else:
        import torch

        def dev(a):
            t = torch.from_numpy(a).cuda()
            return t, t.data_ptr()

        def back(t):
            return t.cpu().numpy()
    B, G = 128, 4
    irs = [orc.synth_ir(9000, c) for c in range(2)]
    n = 90 * B
    x = np.stack([orc.synth_input(2 * n, c) for c in range(2)])
    out = np.zeros_like(x)
    engs = [Engine(2, lib=lib) for _ in range(G)]
    for call in range(2):
        seg = np.ascontiguousarray(x[:, call * n:(call + 1) * n])
        xd, xp = dev(seg)
        for g, e in enumerate(engs):
            if call == 0:
                assert e.init_uniform(B, irs)
            yd, yp = dev(np.zeros_like(seg))
            e.process_device_sliced(xp, n, yp, n, n, g, G, sync=True)
            y = back(yd)
            per = (90 + G - 1) // G
            a, b = min(90, g * per) * B, min(90, (g + 1) * per) * B
            out[:, call * n + a:call * n + b] = y[:, a:b]
            assert np.all(y[:, :a] == 0) and np.all(y[:, b:] == 0)     # nothing outside the slice is touched
    for c in range(2):
        o = orc.OracleUniform()
        o.init(B, irs[c])
        assert peak_err(out[c], o.process(x[c])) <= TOL


def test_sliced_refuses_what_it_cannot_do(lib):
    ir = orc.synth_ir(3000)
    x = orc.synth_input(64 * 10)
    e = Engine(1, lib=lib)
    assert e.init_twostage(16, 256, [ir])
    with pytest.raises(B200ConvError, match="uniform"):
        e.process_sliced([x], [np.empty_like(x)], 0, 2)
    e2 = Engine(1, lib=lib)
    assert e2.init_uniform(64, [ir])
    with pytest.raises(B200ConvError, match="block-aligned"):
        e2.process_sliced([x[:100]], [np.empty(100, np.float32)], 0, 2)
    e2.process([x[:10]])
    with pytest.raises(B200ConvError, match="block-aligned"):
        e2.process_sliced([x], [np.empty_like(x)], 1, 2)
