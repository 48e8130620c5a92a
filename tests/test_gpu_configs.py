"""GPU parity at BASELINE.json's FULL configuration sizes.

* short runs against the CPU oracle (the C restatement; oracle/_ref when present) on the same
  seeded inputs;
* size-independent properties at full length: impulse -> the IR itself, linearity,
  time invariance, chunking invariance between the batched and the streaming kernels.
Tolerance everywhere: 1e-5 of the output peak (north_star).
"""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import Engine
from reevr_b200.synth import synth_input, synth_ir

pytestmark = pytest.mark.gpu
TOL = 1e-5


def peak_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30)


def _oracle_cls(kind):
    if orc.ref_available():
        return orc.RefUniform if kind == "uniform" else orc.RefTwoStage
    return orc.OracleUniform if kind == "uniform" else orc.OracleTwoStage


# (name, C, sr, ir seconds, kind, head/block, tail, blocks to run against the oracle)
CONFIGS = [
    ("cfg1_mono_1s_b512", 1, 48000, 1, "uniform", 512, 0, 200),
    ("metric_stereo_10s_b512", 2, 48000, 10, "uniform", 512, 0, 40),
    ("cfg2_stereo_5s_twostage_b128", 2, 48000, 5, "twostage", 128, 8192, 400),
    ("cfg3_stereo_96k_30s_twostage_b64", 2, 96000, 30, "twostage", 64, 8192, 600),
    ("cfg4_8ch_10s_b512", 8, 48000, 10, "uniform", 512, 0, 12),
    ("cfg5_stereo_120s_b512", 2, 48000, 120, "uniform", 512, 0, 10),
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
def test_full_config_vs_oracle(cfg):
    name, C, sr, secs, kind, blk, tail, nblocks = cfg
    L = sr * secs
    irs = [synth_ir(L, c) for c in range(C)]
    n = nblocks * blk
    xs = [synth_input(n, c) for c in range(C)]
    e = Engine(C)
    assert e.init_uniform(blk, irs) if kind == "uniform" else e.init_twostage(blk, tail, irs)
    # first half in one long call (batched kernels), second half block by block (streaming kernels)
    half = (nblocks // 2) * blk
    ys = [np.empty(n, np.float32) for _ in range(C)]
    for c, y in enumerate(e.process([x[:half] for x in xs])):
        ys[c][:half] = y
    for pos in range(half, n, blk):
        for c, y in enumerate(e.process([x[pos:pos + blk] for x in xs])):
            ys[c][pos:pos + blk] = y
    for c in range(min(C, 3)):       # the oracle is slow; three channels pin the batching
        o = _oracle_cls(kind)()
        assert o.init(blk, irs[c]) if kind == "uniform" else o.init(blk, tail, irs[c])
        yo = o.run(xs[c], blk)
        assert peak_err(ys[c], yo) <= TOL, (name, c)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
def test_impulse_returns_the_ir(cfg):
    """x = a*delta[n-d]  =>  y[n] = a*h[n-d] over the WHOLE IR length (exercises every partition)."""
    name, C, sr, secs, kind, blk, tail, _ = cfg
    if C > 2:
        C = 2
    L = sr * secs
    irs = [synth_ir(L, c) for c in range(C)]
    d, a = 777, 0.5
    n = L + d + 3 * blk
    xs = [np.zeros(n, np.float32) for _ in range(C)]
    for x in xs:
        x[d] = a
    e = Engine(C)
    assert e.init_uniform(blk, irs) if kind == "uniform" else e.init_twostage(blk, tail, irs)
    ys = e.process(xs)
    for c in range(C):
        want = np.zeros(n, np.float32)
        Lt = e.ir_len(c)
        want[d:d + Lt] = a * irs[c][:Lt]
        assert peak_err(ys[c], want) <= TOL, (name, c)


def test_linearity_and_time_invariance_metric_shape():
    L = 480000
    irs = [synth_ir(L, c) for c in range(2)]
    n = 512 * 300
    x1 = [synth_input(n, c) for c in range(2)]
    x2 = [synth_input(n, c + 10) for c in range(2)]
    a, b = 0.7, -1.3

    def run(xs):
        e = Engine(2)
        assert e.init_uniform(512, irs)
        return e.process(xs)

    y1, y2 = run(x1), run(x2)
    y12 = run([a * p + b * q for p, q in zip(x1, x2)])
    for c in range(2):
        assert peak_err(y12[c], a * y1[c] + b * y2[c]) <= TOL
    # shift by a non-multiple of the block size
    s = 1234
    xs = [np.concatenate([np.zeros(s, np.float32), p[:-s]]) for p in x1]
    ysh = run(xs)
    for c in range(2):
        assert np.all(ysh[c][:1024] == 0)           # whole blocks of silence are exactly silent
        assert np.max(np.abs(ysh[c][1024:s])) <= TOL * np.max(np.abs(y1[c]))   # block that contains the onset: rounding only
        assert peak_err(ysh[c][s:], y1[c][:-s]) <= TOL


def test_batch_and_streaming_kernels_agree():
    """The same stream through 1-block calls (k_cmac_stream), ragged calls and one long call (k_cmac_batch)."""
    irs = [synth_ir(480000, c) for c in range(2)]
    n = 512 * 96
    xs = [synth_input(n, c) for c in range(2)]
    outs = []
    for chunks in ([n], [512] * 96, [480] * 102 + [192], [1536] * 32):
        e = Engine(2)
        assert e.init_uniform(512, irs)
        ys = [np.empty(n, np.float32) for _ in range(2)]
        pos = 0
        for k in chunks:
            for c, y in enumerate(e.process([x[pos:pos + k] for x in xs])):
                ys[c][pos:pos + k] = y
            pos += k
        outs.append(ys)
    for other in outs[1:]:
        for c in range(2):
            assert peak_err(other[c], outs[0][c]) <= 2e-6


def test_device_resident_batch_matches_host_path():
    import torch
    irs = [synth_ir(48000, c) for c in range(2)]
    n = 512 * 500
    xs = [synth_input(n, c) for c in range(2)]
    e1 = Engine(2)
    e1.init_uniform(512, irs)
    yh = e1.process(xs)
    e2 = Engine(2)
    e2.init_uniform(512, irs)
    x = torch.from_numpy(np.stack(xs)).cuda()
    y = torch.empty_like(x)
    e2.process_device(x.data_ptr(), n, y.data_ptr(), n, n, sync=True)
    y = y.cpu().numpy()
    for c in range(2):
        assert peak_err(y[c], yh[c]) <= 1e-7


def test_cfg3_true_nonuniform_schedule():
    """Config 3 with a real non-uniform partition schedule (64 / 512 / 4096 / 8192 — beyond the
    reference): same linear convolution, checked against the reference's
    two-stage output on seeded noise and against the impulse identity over the whole 30 s IR."""
    L = 96000 * 30
    irs = [synth_ir(L, c) for c in range(2)]
    blocks, offsets = [64, 512, 4096, 8192], [0, 1024, 8192, 65536]
    e = Engine(2)
    assert e.init_stages(blocks, offsets, irs)
    assert [s["block"] for s in e.stages()] == blocks
    n = 64 * 700
    xs = [synth_input(n, c) for c in range(2)]
    ys = [np.empty(n, np.float32) for _ in range(2)]
    pos = 0
    for k in [64 * 300, 48, 48, 5000, n - 64 * 300 - 96 - 5000]:
        for c, y in enumerate(e.process([x[pos:pos + k] for x in xs])):
            ys[c][pos:pos + k] = y
        pos += k
    o = _oracle_cls("twostage")()
    assert o.init(64, 8192, irs[0])
    assert peak_err(ys[0], o.run(xs[0], 64)) <= TOL
    # impulse -> IR over all four stages
    e2 = Engine(1)
    assert e2.init_stages(blocks, offsets, [irs[1]])
    d = 333
    x = np.zeros(L + d + 100, np.float32)
    x[d] = 1.0
    y = e2.process([x])[0]
    want = np.zeros_like(x)
    want[d:d + e2.ir_len(0)] = irs[1][:e2.ir_len(0)]
    assert peak_err(y, want) <= TOL
