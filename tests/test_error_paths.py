"""Error paths of the C ABI: a recoverable failure must not poison the handle, block sizes above
8192 are accepted (clamped internally, same linear convolution), tile heights that do not divide 32."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as orc
from reevr_b200.convolver import B200ConvError, Engine, TwoStageFFTConvolver
from tests.backends import get_lib, lib  # noqa: F401

TOL = 1e-5


def peak_err(y, ref):
    return float(np.max(np.abs(y - ref)) / max(np.max(np.abs(ref)), 1e-30))


@pytest.mark.parametrize("nth", [0, 3, 7, 11])
def test_out_of_memory_during_init_leaves_the_handle_usable(nth):
    lib_ = get_lib("emu")
    lib_.pc_emu_fail_malloc_after.argtypes = [C.c_int]
    h = orc.synth_ir(3000)
    x = orc.synth_input(64 * 20)
    e = Engine(1, lib=lib_)
    lib_.pc_emu_fail_malloc_after(nth)
    with pytest.raises(B200ConvError) as ei:
        e.init_twostage(16, 256, [h])
    lib_.pc_emu_fail_malloc_after(-1)
    assert "(-4)" in str(ei.value)                       # B200CONV_ENOMEM, not the sticky ECUDA
    assert e.stages() == []                              # nothing half-loaded
    assert np.all(e.process([x[:100]])[0] == 0)          # "no IR" state: zeros, as after reset()
    assert e.init_twostage(16, 256, [h])                 # and the very same handle loads fine afterwards
    o = orc.OracleTwoStage()
    o.init(16, 256, h)
    assert peak_err(e.process([x])[0], o.process(x)) <= TOL


def test_failed_p2p_export_falls_back_to_the_reduce_hook():
    """Injects an allocation failure into the slot-exchange set-up of shard 0, then runs both shards through the
    reduce hook (the documented fallback of distributed.attach_p2p)."""
    lib_ = get_lib("emu")
    lib_.pc_emu_fail_malloc_after.argtypes = [C.c_int]
    h = orc.synth_ir(5000)
    x = orc.synth_input(128 * 30 + 9)
    parked = []
    outs = None
    for rank in (1, 0):
        e = Engine(1, shard_rank=rank, shard_count=2, lib=lib_)
        assert e.init_uniform(128, [h])
        if rank == 0:
            lib_.pc_emu_fail_malloc_after(1)
            with pytest.raises(B200ConvError):
                e.p2p_export(mode=1)
            lib_.pc_emu_fail_malloc_after(-1)

        def hook(ptr, n, stream, rank=rank):
            a = np.ctypeslib.as_array((C.c_float * n).from_address(ptr))
            if rank == 1:
                parked.append(a.copy())
            else:
                a += parked.pop(0)
            return 0
        e.set_reduce(hook)
        ys = [e.process([c])[0] for c in (x[:2000], x[2000:])]
        if rank == 0:
            outs = np.concatenate(ys)
    o = orc.OracleUniform()
    o.init(128, h)
    assert peak_err(outs, o.process(x)) <= TOL


def test_block_sizes_above_8192_are_clamped_not_rejected(lib):
    """REEV-R: tail = max(8192, 2*head) (StereoConvolver.cpp:15) is 16384 for host blocks above 4096 samples;
    the reference's own two-stage perf shape is 100/16384.  Same linear convolution, any partition size."""
    h = orc.synth_ir(70000)
    x = orc.synth_input(60000)
    t = TwoStageFFTConvolver(lib=lib)
    assert t.init(100, 16384, h)
    o = orc.OracleTwoStage()
    o.init(100, 16384, h)
    y = np.concatenate([t.process(x[i:i + 4100]) for i in range(0, x.size, 4100)])
    assert peak_err(y, o.process(x)) <= TOL
    e = Engine(1, lib=lib)
    assert e.init_uniform(20000, [h[:40000]])            # uniform 32768 -> 8192 internally
    assert e.stages()[0]["block"] == 8192
    o2 = orc.OracleUniform()
    o2.init(32768, h[:40000])
    assert peak_err(e.process([x])[0], o2.process(x)) <= TOL


@pytest.mark.parametrize("variant", [25, 28, 16, 6])
@pytest.mark.parametrize("nparts", [25, 32, 47, 97])
def test_tile_heights_that_do_not_divide_the_row_padding(lib, variant, nparts):
    """TT = 12 / 24 / 32 with partition counts whose round-up to TT exceeds the next multiple of 32, several
    channels (the rows behind channel c's H are channel c+1's, which are not zero)."""
    B = 32
    irs = [orc.synth_ir(nparts * B - 5, c) for c in range(3)]
    x = [orc.synth_input(B * 70, c) for c in range(3)]
    e = Engine(3, cmac_variant=variant, lib=lib)
    assert e.init_uniform(B, irs)
    ys = e.process(x)
    for c in range(3):
        o = orc.OracleUniform()
        o.init(B, irs[c])
        assert peak_err(ys[c], o.process(x[c])) <= TOL
