"""Tensor-core form of the batched sweep (reevr_b200/csrc/kernels_tc.cuh, cmac_variant 40): tcgen05 kind::tf32 with the
3xTF32 split, block-Toeplitz tiles of H against row-shifted windows of the per-bin time lines.  GPU only (the CPU
emulation has no tensor memory); checked against the oracle (FFTConvolver.cpp:176-187 restated) at the usual 1e-5 of
peak and against the FFMA sweep of the same engine at a tighter bound, over ragged launch groups, the largest
supported partition count, a single partition, the DC / Nyquist entry and multi-stage handles."""
import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import B200ConvError, Engine
from tests.backends import get_lib

pytestmark = pytest.mark.gpu
TOL = 1e-5


def peak_err(y, ref):
    y = np.asarray(y, np.float64)
    ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


def run(eng, xs, chunks):
    outs = [[] for _ in xs]
    pos = 0
    for k in chunks:
        ys = eng.process([x[pos:pos + k] for x in xs])
        for c, y in enumerate(ys):
            outs[c].append(y)
        pos += k
    return [np.concatenate(o) for o in outs]


@pytest.mark.parametrize("B,nparts,nblocks,C", [(64, 100, 700, 2), (512, 938, 260, 2), (32, 961, 150, 1), (128, 1, 300, 2),
                                                (256, 65, 1100, 4)])
def test_tc_sweep_matches_oracle_and_ffma(B, nparts, nblocks, C):
    lib = get_lib("cuda")
    irs = [orc.synth_ir(nparts * B - (5 if nparts > 1 else 0), c) for c in range(C)]
    n = nblocks * B + 37
    xs = [orc.synth_input(n, c) for c in range(C)]
    chunks = [n // 3 + 11, B - 11, n - (n // 3 + 11) - (B - 11)]       # ragged launch groups, an open block in between
    ys = {}
    for variant in (40, 22):
        e = Engine(C, cmac_variant=variant, max_batch_blocks=512, lib=lib)
        assert e.init_uniform(B, irs)
        ys[variant] = run(e, xs, chunks)
        e.close()
    for c in range(C):
        o = orc.OracleUniform()
        o.init(B, irs[c])
        ref = o.process(xs[c])
        assert peak_err(ys[40][c], ref) <= TOL
        assert peak_err(ys[40][c], ys[22][c]) <= 4e-6


def test_tc_sweep_is_the_default_for_long_launch_groups():
    import torch
    lib = get_lib("cuda")
    B, nparts, T = 64, 100, 4608
    irs = [orc.synth_ir(nparts * B - 9, c) for c in range(2)]
    n = T * B
    x = np.stack([orc.synth_input(n, c) for c in range(2)])
    xd = torch.from_numpy(x).cuda()
    res, launches = {}, {}
    for tc in (1, 0):
        e = Engine(2, max_batch_blocks=T + 1, lib=lib)
        assert e.init_uniform(B, irs)
        e.set_option("tc", tc)
        yd = torch.zeros_like(xd)
        before = e.launch_count
        e.process_device(xd.data_ptr(), n, yd.data_ptr(), n, n, sync=True)
        launches[tc] = e.launch_count - before
        assert e.last_sweep_variant() == (40 if tc else 22)
        res[tc] = yd.cpu().numpy()
        e.close()
    assert launches[1] == launches[0] + 3              # Toeplitz images + time lines + merge on top of the sweep itself
    for c in range(2):
        assert peak_err(res[1][c], res[0][c]) <= 4e-6
    o = orc.OracleUniform()
    o.init(B, irs[0])
    assert peak_err(res[1][0][:600 * B], o.process(x[0][:600 * B])) <= TOL


def test_tc_sweep_twostage_handle():
    lib = get_lib("cuda")
    irs = [orc.synth_ir(60000, c) for c in range(2)]
    n = 128 * 900 + 50
    xs = [orc.synth_input(n, c) for c in range(2)]
    e = Engine(2, cmac_variant=40, max_batch_blocks=400, lib=lib)
    assert e.init_twostage(128, 8192, irs)
    ys = run(e, xs, [n // 2 + 3, n - n // 2 - 3])
    e.close()
    for c in range(2):
        o = orc.OracleTwoStage()
        o.init(128, 8192, irs[c])
        assert peak_err(ys[c], o.process(xs[c])) <= TOL


def test_tc_sweep_refuses_what_it_cannot_do():
    lib = get_lib("cuda")
    B = 32
    irs = [orc.synth_ir(962 * B, 0)]                    # 962 partitions: one more than a tile's K range covers
    x = [orc.synth_input(40 * B, 0)]
    e = Engine(1, cmac_variant=40, lib=lib)
    assert e.init_uniform(B, irs)
    with pytest.raises(B200ConvError):
        e.process(x)
    e.close()
    e = Engine(1, lib=lib)                               # automatic selection never picks it for that shape
    assert e.init_uniform(B, irs)
    y = e.process(x)[0]
    o = orc.OracleUniform()
    o.init(B, irs[0])
    assert peak_err(y, o.process(x[0])) <= TOL
    e.close()
