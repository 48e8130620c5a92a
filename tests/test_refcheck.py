"""The windowed / IR-segmented reference shortcuts bench.py's multi-GPU parity objects rely on reproduce the plain
reference run (oracle/refcheck.py)."""
import numpy as np

from oracle import oracle as orc
from oracle import refcheck as rc


def test_window_and_segmented_window_equal_the_full_reference_run():
    B = 64
    ir = orc.synth_ir(57 * B - 11)
    x = orc.synth_input(300 * B)
    o = orc.OracleUniform()
    o.init(B, ir)
    full = o.run(x, B)
    for w0 in (0, 3, 58, 200, 236):
        want = full[w0 * B:(w0 + 64) * B]
        got = rc.ref_window(B, ir, x, w0, 64)
        assert rc.peak_err(got, want) <= 2e-7
        for nseg in (1, 4, 7, 57):
            seg = rc.ref_window_segmented(B, ir, x, w0, 64, nseg=nseg, threads=4)
            assert rc.peak_err(seg, want) <= 1e-6, (w0, nseg)


def test_trim_rule():
    ir = np.array([1, 0.5, 1e-7, 0, -2e-7], np.float32)
    assert rc.trimmed(ir).size == 2
    assert rc.ref_window(8, np.zeros(5, np.float32), np.ones(64, np.float32), 2, 2).tolist() == [0.0] * 16
