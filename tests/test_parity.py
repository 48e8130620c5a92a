"""Parity of the engine (through the C ABI) with the CPU oracle.

Every test runs twice: on the CPU emulation of the kernels ("emu", no GPU needed) and on the
real CUDA library ("cuda", marked gpu).  Tolerance: max |y - y_oracle| / max |y_oracle| <= 1e-5
(BASELINE.json north_star: "<= 1e-5 max relative error vs CPU reference", normalised by the
output peak as SURVEY §8c defines it).
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import Engine, FFTConvolver, StereoConvolver, TwoStageFFTConvolver
from tests import refcases as rc
from tests.backends import lib  # noqa: F401  (fixture)
from tests.golden.make_golden import run_case

TOL = 1e-5
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def peak_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def ragged_chunks(n, hi, seed):
    rng = np.random.default_rng(seed)
    out, done = [], 0
    while done < n:
        k = int(min(n - done, rng.integers(1, hi)))
        out.append(k)
        done += k
    return out


def run_chunks(conv, x, chunks):
    y = np.empty_like(x)
    pos = 0
    for k in chunks:
        y[pos:pos + k] = conv.process(x[pos:pos + k])
        pos += k
    return y


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))), ids=os.path.basename)
def test_golden_fixtures(lib, path):
    g = np.load(path, allow_pickle=False)
    spec = {k: (g[k].item() if g[k].ndim == 0 else g[k]) for k in g.files if k != "out"}
    factory = lambda kind: FFTConvolver(lib=lib) if kind == "uniform" else TwoStageFFTConvolver(lib=lib)
    y = run_case(spec, impl=factory)
    assert peak_err(y, g["out"]) <= TOL


@pytest.mark.parametrize("case", rc.UNIFORM_CASES, ids=lambda c: "u-" + "-".join(map(str, c)))
def test_reference_selftest_uniform(lib, case):
    n_in, n_ir, bmin, bmax, block = case
    x, h = rc.ramp(n_in), rc.ramp(n_ir)
    total = n_in + n_ir - 1
    chunks = rc.chunk_schedule(total, bmin, bmax, rc.GlibcRand(1))
    conv = FFTConvolver(lib=lib)
    assert conv.init(block, h)
    y = rc.drive(conv, x, total, chunks)
    o = orc.OracleUniform()
    o.init(block, h)
    yo = rc.drive(o, x, total, chunks)
    assert peak_err(y, yo) <= TOL
    assert rc.reference_selftest_ok(y, orc.naive_convolve(x, h), n_ir)


@pytest.mark.parametrize("case", rc.TWOSTAGE_CASES, ids=lambda c: "t-" + "-".join(map(str, c)))
def test_reference_selftest_twostage(lib, case):
    n_in, n_ir, bmin, bmax, head, tail = case
    x, h = rc.ramp(n_in), rc.ramp(n_ir)
    total = n_in + n_ir - 1
    chunks = rc.chunk_schedule(total, bmin, bmax, rc.GlibcRand(1))
    conv = TwoStageFFTConvolver(lib=lib)
    assert conv.init(head, tail, h)
    y = rc.drive(conv, x, total, chunks)
    o = orc.OracleTwoStage()
    o.init(head, tail, h)
    yo = rc.drive(o, x, total, chunks)
    assert peak_err(y, yo) <= TOL
    assert rc.reference_selftest_ok(y, orc.naive_convolve(x, h), n_ir)


def test_error_conventions(lib):
    # FFTConvolver.cpp:97-111,157-161
    c = FFTConvolver(lib=lib)
    assert not c.init(0, np.ones(4, np.float32))
    assert np.all(c.process(np.ones(7, np.float32)) == 0)         # process before a successful init
    assert c.init(8, np.zeros(0, np.float32))
    assert np.all(c.process(np.ones(5, np.float32)) == 0)
    assert c.init(8, np.full(16, 5e-7, np.float32))               # everything below the trim threshold
    assert np.all(c.process(np.ones(5, np.float32)) == 0)
    h = np.ones(20, np.float32)
    h[17:] = 9e-7
    assert c.init(5, h)
    st = c._e.stages()
    assert st[0]["block"] == 8 and st[0]["partitions"] == 3 and c._e.ir_len(0) == 17
    t = TwoStageFFTConvolver(lib=lib)
    assert not t.init(0, 8, h) and not t.init(8, 0, h)
    assert t.init(16, 4, h)          # head > tail is swapped (TwoStageFFTConvolver.cpp:100-104)
    assert t._e.stages()[0]["block"] == 4
    assert c.process(np.zeros(0, np.float32)).size == 0


@pytest.mark.parametrize("block,ir_len", [(64, 1000), (16, 333), (256, 5000), (1, 9), (2, 31)])
def test_chunking_invariance_and_oracle(lib, block, ir_len):
    h = orc.synth_ir(ir_len)
    n = 40 * max(block, 8)
    x = orc.synth_input(n)
    o = orc.OracleUniform()
    o.init(block, h)
    yo = o.process(x)
    bp = o.block
    a = FFTConvolver(lib=lib)
    a.init(block, h)
    y_whole = a.process(x)                                        # one long call
    b = FFTConvolver(lib=lib)
    b.init(block, h)
    y_blocks = run_chunks(b, x, [bp] * (n // bp))                  # block by block
    c = FFTConvolver(lib=lib)
    c.init(block, h)
    y_rag = run_chunks(c, x, ragged_chunks(n, 3 * bp + 2, seed=block))
    assert peak_err(y_whole, yo) <= TOL
    assert peak_err(y_blocks, yo) <= TOL
    assert peak_err(y_rag, yo) <= TOL
    assert peak_err(y_rag, y_whole) <= 2e-6


@pytest.mark.parametrize("head,tail,ir_len", [(32, 256, 3000), (8, 8, 100), (64, 128, 100), (16, 64, 129), (4, 1024, 5000)])
def test_twostage_vs_oracle(lib, head, tail, ir_len):
    h = orc.synth_ir(ir_len)
    n = 6 * tail + 37
    x = orc.synth_input(n)
    o = orc.OracleTwoStage()
    o.init(head, tail, h)
    yo = o.process(x)
    for chunks in ([n], ragged_chunks(n, 3 * head + 2, seed=tail), [head] * (n // head) + [n % head]):
        chunks = [k for k in chunks if k]
        t = TwoStageFFTConvolver(lib=lib)
        assert t.init(head, tail, h)
        assert peak_err(run_chunks(t, x, chunks), yo) <= TOL


def test_clear_is_a_true_clear(lib):
    # block aligned: identical to the reference (zeros afterwards); mid-block: the engine still
    # silences everything (documented deviation from the quirk of FFTConvolver.cpp:80-90)
    h = orc.synth_ir(1000)
    x = orc.synth_input(1024)
    for conv, init in ((FFTConvolver(lib=lib), lambda c: c.init(64, h)),
                       (TwoStageFFTConvolver(lib=lib), lambda c: c.init(32, 128, h))):
        init(conv)
        for upto in (512, 500):
            conv.process(x[:upto])
            conv.clear()
            assert np.all(conv.process(np.zeros(2048, np.float32)) == 0)
        # after a clear the convolver behaves like a fresh one
        conv.clear()
        fresh = type(conv)(lib=lib)
        init(fresh)
        assert peak_err(conv.process(x), fresh.process(x)) <= 1e-7


def test_reset_and_reinit(lib):
    c = FFTConvolver(lib=lib)
    h1, h2 = orc.synth_ir(700, 0), orc.synth_ir(1500, 1)
    x = orc.synth_input(2048)
    c.init(128, h1)
    c.process(x)
    c.reset()
    assert np.all(c.process(x[:100]) == 0)                        # IR dropped (FFTConvolver.cpp:56-78)
    assert c.init(32, h2)                                         # re-init with another shape
    o = orc.OracleUniform()
    o.init(32, h2)
    assert peak_err(c.process(x), o.process(x)) <= TOL
    assert c.init(64, h1)                                         # init() on a live object resets first (:95)
    o.init(64, h1)
    assert peak_err(c.process(x), o.process(x)) <= TOL


@pytest.mark.parametrize("C", [2, 4, 8])
def test_multichannel_engine(lib, C):
    # channels are independent convolvers with their own (differently long) IRs — config 4's shape, small
    irs = [orc.synth_ir(900 + 157 * c, c) for c in range(C)]
    xs = [orc.synth_input(3000, c) for c in range(C)]
    e = Engine(C, lib=lib)
    assert e.init_uniform(64, irs)
    ys = e.process(xs)
    for c in range(C):
        o = orc.OracleUniform()
        o.init(64, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL


def test_stereo_convolver_quad(lib):
    # src/dsp/StereoConvolver.cpp:8-42: head = nextPow2(host block), tail = max(8192, 2*head)
    sc = StereoConvolver(lib=lib)
    sc.prepare(48)
    assert sc.headBlockSize == 64 and sc.tailBlockSize == 8192
    irs = [orc.synth_ir(20000, c) for c in range(4)]
    sc.loadImpulse(*irs)
    L, R = orc.synth_input(48 * 50, 0), orc.synth_input(48 * 50, 1)
    outs = [np.empty_like(L) for _ in range(4)]
    for i in range(50):
        seg = slice(48 * i, 48 * (i + 1))
        for buf, y in zip(outs, sc.process(L[seg], R[seg])):
            buf[seg] = y
    for c, (buf, src) in enumerate(zip(outs, (L, R, L, R))):
        o = orc.OracleTwoStage()
        o.init(64, 8192, irs[c])
        assert peak_err(buf, run_chunks(o, src, [48] * 50)) <= TOL


def test_nonuniform_stages(lib):
    # a 3-stage schedule (beyond the reference) is the same linear convolution
    h = orc.synth_ir(6000)
    x = orc.synth_input(5000)
    o = orc.OracleUniform()
    o.init(256, h)
    yo = o.process(x)
    e = Engine(1, lib=lib)
    assert e.init_stages([16, 64, 512], [0, 128, 1024], [h])
    st = e.stages()
    assert [s["block"] for s in st] == [16, 64, 512]
    y = np.concatenate([e.process([x[i:i + 777]])[0] for i in range(0, 5000, 777)])
    assert peak_err(y, yo) <= TOL


def test_long_call_is_split_into_groups(lib):
    # max_batch_blocks small -> a long call runs as several launch groups + timeline compaction
    h = orc.synth_ir(2000)
    x = orc.synth_input(64 * 300)
    e = Engine(1, max_batch_blocks=16, lib=lib)
    e.init_uniform(64, [h])
    y = e.process([x])[0]
    o = orc.OracleUniform()
    o.init(64, h)
    assert peak_err(y, o.run(x, 64)) <= TOL


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 11, 12, 16, 21, 22, 23, 24, 25, 26, 27, 28, 33, 34])
def test_cmac_variants(lib, variant):
    h = orc.synth_ir(3000)
    x = orc.synth_input(64 * 100)
    e = Engine(1, cmac_variant=variant, lib=lib)
    e.init_uniform(64, [h])
    y = e.process([x])[0]
    o = orc.OracleUniform()
    o.init(64, h)
    assert peak_err(y, o.run(x, 64)) <= TOL


_STREAM_CASES = [(v, B, n) for v in (100, 101, 102, 103, 104, 105, 106, 107, 108)
                 for B, n in ((64, 37), (128, 9), (512, 21), (1024, 5), (2048, 3))
                 if not (v == 100 and B > 512)]        # the generic fallback kernel is only ever selected below 64 bins


@pytest.mark.parametrize("variant,B,nparts", _STREAM_CASES)
def test_streaming_sweep_variants(lib, variant, B, nparts):
    """One block per launch (the real-time call): register-batch and TMA-ring forms of the streaming sweep, block
    sizes on both sides of the 512-bin CTA tile, partition counts that leave ragged ring stages, 2 channels."""
    irs = [orc.synth_ir(nparts * B - 3, c) for c in range(2)]
    x = [orc.synth_input(B * 12 + 40, c) for c in range(2)]
    e = Engine(2, cmac_variant=variant, lib=lib)
    assert e.init_uniform(B, irs)
    chunks = [B] * 6 + [40, B - 40] + [B] * 5            # whole blocks, one block in two calls
    ys = [[], []]
    pos = 0
    for k in chunks:
        out = e.process([a[pos:pos + k] for a in x])
        for c in range(2):
            ys[c].append(out[c])
        pos += k
    for c in range(2):
        o = orc.OracleUniform()
        o.init(B, irs[c])
        assert peak_err(np.concatenate(ys[c]), o.process(x[c][:pos])) <= TOL


def test_device_mixdown_quad(lib):
    """StereoConvolver quad as ONE call with 2 inputs / 2 outputs and the true-stereo
    mixdown on the device (src/PluginProcessor.cpp:1833-1838: wet L = LL + RL, wet R = RR + LR)."""
    sc = StereoConvolver(lib=lib)
    sc.prepare(100)                                    # head 128, tail 8192
    irs = [orc.synth_ir(20000, c) for c in range(4)]   # LL, RR, LR, RL
    sc.loadImpulse(*irs)
    sc.enable_device_mixdown(true_stereo=True)
    n = 100 * 40
    L, R = orc.synth_input(n, 0), orc.synth_input(n, 1)
    wl, wr = np.empty_like(L), np.empty_like(R)
    for i in range(40):
        seg = slice(100 * i, 100 * (i + 1))
        wl[seg], wr[seg] = sc.process_mixed(L[seg], R[seg])
    outs = []
    for ir, src in zip(irs, (L, R, L, R)):
        o = orc.OracleTwoStage()
        o.init(128, 8192, ir)
        outs.append(run_chunks(o, src, [100] * 40))
    LL, RR, LR, RL = outs
    assert peak_err(wl, LL + RL) <= TOL
    assert peak_err(wr, RR + LR) <= TOL
    # long call through the pipelined path + removing the routing again
    e = Engine(4, lib=lib, max_batch_blocks=32)
    assert e.init_uniform(64, [ir[:3000] for ir in irs])
    e.set_routing([0, 1, 0, 1], [[1, 0, 0, 1], [0, 1, 1, 0]])
    yl, yr = e.process([L, R])
    refs = []
    for ir, src in zip(irs, (L, R, L, R)):
        o = orc.OracleUniform()
        o.init(64, ir[:3000])
        refs.append(o.process(src))
    assert peak_err(yl, refs[0] + refs[3]) <= TOL and peak_err(yr, refs[1] + refs[2]) <= TOL
    e.set_routing(None, None)
    assert len(e.process([L[:64], R[:64], L[:64], R[:64]])) == 4


def test_ir_hot_swap_prime_and_crossfade(lib):
    """warm the incoming convolver with the last 0.25 s in ONE call (prime), then crossfade
    old/new for 50 ms on the device — the sequence of src/PluginProcessor.cpp:1695-1750,1800-1830."""
    sr, blk = 48000, 128
    h_old = [orc.synth_ir(30000, c) for c in range(2)]
    h_new = [orc.synth_ir(24000, c + 7) for c in range(2)]
    n_hist, n_fade, n_after = sr // 4, 2400, 4000
    x = [orc.synth_input(20000 + n_fade + n_after, c) for c in range(2)]
    old = Engine(2, lib=lib)
    new = Engine(2, lib=lib)
    assert old.init_twostage(blk, 8192, h_old) and new.init_twostage(blk, 8192, h_new)
    pre = [a[:20000] for a in x]
    y_pre = old.process(pre)
    new.prime([a[20000 - n_hist:20000] for a in x])                      # warm-up replay, no output
    fade_in = [a[20000:20000 + n_fade] for a in x]
    y_fade = old.process_xfade(new, fade_in, 0.0, 1.0 / n_fade)
    y_post = new.process([a[20000 + n_fade:] for a in x])
    for c in range(2):
        oo, on = orc.OracleTwoStage(), orc.OracleTwoStage()
        oo.init(blk, 8192, h_old[c])
        on.init(blk, 8192, h_new[c])
        r_pre = oo.process(pre[c])
        on.process(x[c][20000 - n_hist:20000])
        a = np.clip(np.arange(n_fade, dtype=np.float32) / n_fade, 0, 1)
        r_fade = (1 - a) * oo.process(fade_in[c]) + a * on.process(fade_in[c])
        r_post = on.process(x[c][20000 + n_fade:])
        assert peak_err(y_pre[c], r_pre) <= TOL
        assert peak_err(y_fade[c], r_fade) <= TOL
        assert peak_err(y_post[c], r_post) <= TOL


def test_ir_decay_eq_stft(lib):
    """SURVEY 8f-3 groundwork: device STFT decay-EQ (Impulse::applyDecay) against the C restatement."""
    from reevr_b200.convolver import ir_decay_eq
    sr = 48000.0
    for n in (30000, 4096, 5000, 1):
        h = orc.synth_ir(n)
        for lut in (np.ones(2049), np.linspace(1.0, 0.7, 2049), np.linspace(0.8, 1.05, 2049)):
            want = orc.apply_decay(h, lut, sr)
            got = ir_decay_eq(h, lut, sr, lib=lib)
            peak = max(np.max(np.abs(want)), 1e-30)
            # the first hop is ill-conditioned in the reference itself (window starts at 0)
            assert np.max(np.abs(got[1024:] - want[1024:])) <= 1e-5 * peak if n > 1024 else True
            assert np.max(np.abs(got[64:1024] - want[64:1024])) <= 2e-3 * peak if n > 64 else True
