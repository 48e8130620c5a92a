"""Kernel-variant sweep on one GPU: per-variant CUDA-event time of the FDL sweep (k_cmac_batch*),
forward FFT and inverse FFT at the metric shape.  python tools/sweep.py [--blocks T ...] [--variants v ...]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reevr_b200.convolver import Engine  # noqa: E402
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, nargs="+", default=[4736, 7104])
ap.add_argument("--variants", type=int, nargs="+", default=[1, 2, 11, 12, 16, 21, 22, 23, 24, 25, 26, 27, 28])
ap.add_argument("--ir_s", type=int, default=10)
ap.add_argument("--block", type=int, default=512)
ap.add_argument("--channels", type=int, default=2)
args = ap.parse_args()

C, B = args.channels, args.block
irs = [synth_ir(args.ir_s * 48000, c) for c in range(C)]
for T in args.blocks:
    n = T * B
    x = torch.from_numpy(np.stack([synth_input(n, c) for c in range(C)])).cuda()
    y = torch.empty_like(x)
    ref = None
    for v in args.variants:
        e = Engine(C, max_batch_blocks=T + 1, cmac_variant=v)
        e.init_uniform(B, irs)
        P = e.stages()[0]["partitions"]
        e.set_timing(True)
        best = None
        for _ in range(4):
            e.process_device(x.data_ptr(), n, y.data_ptr(), n, n, sync=True)
            t = e.last_timing()
            if best is None or t["cmac_ms"] < best["cmac_ms"]:
                best = t
        out = y.clone()
        if ref is None:
            ref = out
        err = float((out - ref).abs().max() / ref.abs().max())
        tf = 8.0 * P * B * C * T / (best["cmac_ms"] * 1e-3) / 1e12
        print(f"T={T:6d} variant {v:3d}: cmac {best['cmac_ms']:.3f} ms ({tf:5.1f} TFLOP/s)  fft {best['fft_ms']:.3f}  ifft {best['ifft_ms']:.3f}"
              f"  -> {n / (best['cmac_ms'] + best['fft_ms'] + best['ifft_ms']) / 1e3:7.1f} M frames/s   maxdiff vs first {err:.1e}", flush=True)
        e.close()
