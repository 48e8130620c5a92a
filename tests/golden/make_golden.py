"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference or a prebuilt oracle/_ref):
    python -m tests.golden.make_golden
Each fixture stores the case parameters, a checksum of the regenerated inputs and the
reference output `out`; `run_case` re-creates the inputs and drives any implementation that
offers the reference surface (init/process/clear), so the same function checks the C oracle
(tests/test_oracle.py) and the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402
from tests import refcases as rc  # noqa: E402

# name -> spec.  kind: uniform|twostage; signal: synth|ramp; chunking: fixed n | ragged(seed)
CASES = {
    "uniform_b64_ir1000": dict(kind="uniform", block=64, tail=0, ir_len=1000, n=4096, signal="synth", chunk=64, rag=0),
    "uniform_b64_ir1000_ragged": dict(kind="uniform", block=64, tail=0, ir_len=1000, n=4096, signal="synth", chunk=0, rag=7),
    "uniform_b512_ir48000_cfg1": dict(kind="uniform", block=512, tail=0, ir_len=48000, n=512 * 24, signal="synth", chunk=512, rag=0),
    "uniform_b512_ir480000_metric": dict(kind="uniform", block=512, tail=0, ir_len=480000, n=512 * 16, signal="synth", chunk=512, rag=0),
    "uniform_b100_ramp": dict(kind="uniform", block=100, tail=0, ir_len=321, n=3000, signal="ramp", chunk=0, rag=3),
    "uniform_b1_tiny": dict(kind="uniform", block=1, tail=0, ir_len=5, n=40, signal="synth", chunk=0, rag=11),
    "twostage_h32_t256_ir3000": dict(kind="twostage", block=32, tail=256, ir_len=3000, n=4096, signal="synth", chunk=32, rag=0),
    "twostage_h32_t256_ir3000_ragged": dict(kind="twostage", block=32, tail=256, ir_len=3000, n=4096, signal="synth", chunk=0, rag=5),
    "twostage_h128_t8192_ir240000_cfg2": dict(kind="twostage", block=128, tail=8192, ir_len=240000, n=128 * 160, signal="synth", chunk=128, rag=0),
    "twostage_h64_t128_short_ir": dict(kind="twostage", block=64, tail=128, ir_len=100, n=1000, signal="synth", chunk=0, rag=9),
    "uniform_clear_midstream": dict(kind="uniform", block=64, tail=0, ir_len=1000, n=4096, signal="synth", chunk=64, rag=0, clear_at=2048),
    "twostage_clear_midstream": dict(kind="twostage", block=32, tail=256, ir_len=3000, n=4096, signal="synth", chunk=32, rag=0, clear_at=2048),
}


def _signals(spec):
    n, L = int(spec["n"]), int(spec["ir_len"])
    if str(spec["signal"]) == "ramp":
        return rc.ramp(n), rc.ramp(L)
    return orc.synth_input(n), orc.synth_ir(L)


def _chunks(spec):
    n = int(spec["n"])
    if int(spec["rag"]) == 0:
        c = int(spec["chunk"])
        return [c] * (n // c) + ([n % c] if n % c else [])
    rng = np.random.default_rng(int(spec["rag"]))
    hi = 3 * int(spec["block"]) + 2
    out, done = [], 0
    while done < n:
        k = int(min(n - done, rng.integers(1, hi)))
        out.append(k)
        done += k
    return out


def make_impl(spec, impl):
    kind = str(spec["kind"])
    if impl == "ref":
        return orc.RefUniform() if kind == "uniform" else orc.RefTwoStage()
    if impl == "oracle":
        return orc.OracleUniform() if kind == "uniform" else orc.OracleTwoStage()
    return impl(kind)  # factory supplied by the caller (CUDA path)


def run_case(spec, impl="oracle"):
    x, h = _signals(spec)
    conv = make_impl(spec, impl)
    if str(spec["kind"]) == "uniform":
        assert conv.init(int(spec["block"]), h)
    else:
        assert conv.init(int(spec["block"]), int(spec["tail"]), h)
    if "in_crc" in spec:
        assert zlib.crc32(x.tobytes()) == int(spec["in_crc"]), "synthetic input generator drifted"
        assert zlib.crc32(h.tobytes()) == int(spec["ir_crc"]), "synthetic IR generator drifted"
    clear_at = int(spec.get("clear_at", -1))
    y = np.empty_like(x)
    pos = 0
    for k in _chunks(spec):
        if pos == clear_at:   # block-aligned clear
            conv.clear()
        y[pos:pos + k] = conv.process(x[pos:pos + k])
        pos += k
    return y


def main():
    assert orc.ref_available(), "needs oracle/_ref (the compiled reference)"
    for name, spec in CASES.items():
        x, h = _signals(spec)
        spec = dict(spec, in_crc=zlib.crc32(x.tobytes()), ir_crc=zlib.crc32(h.tobytes()))
        out = run_case(spec, impl="ref")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), out=out, **spec)
        print(f"{name}: {out.size} samples, peak {np.abs(out).max():.4g}")


if __name__ == "__main__":
    main()
