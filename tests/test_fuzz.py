"""Randomised schedule / chunking fuzz of the engine against the oracle (CPU emulation + GPU).

Every case draws: block size(s), IR length, channel count, a stage schedule (uniform / two-stage /
3-stage non-uniform), a small launch-group size (forces group splitting + timeline compaction),
random call lengths (1 .. several blocks), one mid-stream clear() on a block boundary, and checks
the whole output against the C oracle (uniform or two-stage reference of the same IR)."""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import Engine
from tests.backends import lib  # noqa: F401

TOL = 1e-5


def _case(seed):
    rng = np.random.default_rng(seed)
    kind = ["uniform", "twostage", "stages"][seed % 3]
    C = int(rng.integers(1, 4))
    if kind == "uniform":
        B = int(2 ** rng.integers(0, 8))
        L = int(rng.integers(1, 40 * B + 2))
    elif kind == "twostage":
        B = int(2 ** rng.integers(0, 6))
        T = B * int(2 ** rng.integers(0, 4))
        L = int(rng.integers(1, 7 * T + 2))
    else:
        B = int(2 ** rng.integers(1, 5))
        L = int(rng.integers(200 * B, 400 * B))
    n = int(rng.integers(20 * B, 120 * B)) + int(rng.integers(0, B))
    return rng, kind, C, B, L, n


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_schedules_and_chunking(lib, seed):
    rng, kind, C, B, L, n = _case(seed)
    irs = [orc.synth_ir(L, c + seed) for c in range(C)]
    xs = [orc.synth_input(n, c + seed) for c in range(C)]
    e = Engine(C, max_batch_blocks=int(rng.integers(3, 40)), lib=lib)
    if kind == "uniform":
        assert e.init_uniform(B, irs)
        mk = lambda ir: (orc.OracleUniform(), (B, ir))
    elif kind == "twostage":
        T = B * int(2 ** rng.integers(0, 4))
        assert e.init_twostage(B, T, irs)
        mk = lambda ir: (orc.OracleTwoStage(), (B, T, ir))
    else:
        blocks = [B, 4 * B, 16 * B]
        offsets = [0, 8 * B, 64 * B]
        assert e.init_stages(blocks, offsets, irs)
        mk = lambda ir: (orc.OracleUniform(), (4 * B, ir))
    # chunk schedule with one block-aligned clear (aligned to the largest block size in play)
    big = 16 * B if kind == "stages" else (B * 16 if kind == "twostage" else B)
    clear_at = (n // 2) // big * big
    chunks, pos = [], 0
    while pos < n:
        k = int(min(n - pos, rng.integers(1, 5 * B + 2)))
        if pos < clear_at < pos + k:
            k = clear_at - pos
        chunks.append(k)
        pos += k
    ys = [np.empty(n, np.float32) for _ in range(C)]
    oracles = []
    for c in range(C):
        o, a = mk(irs[c])
        assert o.init(*a)
        oracles.append(o)
    refs = [np.empty(n, np.float32) for _ in range(C)]
    pos = 0
    for k in chunks:
        if pos == clear_at and pos > 0:
            e.clear()
            for o in oracles:
                o.clear()
        out = e.process([x[pos:pos + k] for x in xs])
        for c in range(C):
            ys[c][pos:pos + k] = out[c]
            refs[c][pos:pos + k] = oracles[c].process(xs[c][pos:pos + k])
        pos += k
    for c in range(C):
        peak = max(np.max(np.abs(refs[c])), 1e-30)
        assert np.max(np.abs(ys[c] - refs[c])) / peak <= TOL, (seed, kind, C, B, L, n, c)
