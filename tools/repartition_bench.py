"""Offline rendering is free to choose its partition size (the output is the same linear convolution): times the
metric job (stereo, 10 s IR, 57.7 M frames) through uniform handles of block 512 (the metric's configuration), 2048,
4096 and 8192 on one GPU and checks that the outputs agree."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reevr_b200.convolver import Engine  # noqa: E402
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402

C, L = 2, 480000
n = 112608 * 512
irs = [synth_ir(L, c) for c in range(C)]
x = torch.from_numpy(np.stack([synth_input(n, c) for c in range(C)])).cuda()
ref = None
for B in (512, 2048, 4096, 8192):
    T = n // B
    e = Engine(C, max_batch_blocks=T + 1)
    assert e.init_uniform(B, irs)
    y = torch.empty_like(x)
    e.set_timing(True)
    best = None
    for _ in range(3):
        e.process_device(x.data_ptr(), n, y.data_ptr(), n, n, sync=True)
        t = e.last_timing()
        tot = t["cmac_ms"] + t["fft_ms"] + t["ifft_ms"]
        if best is None or tot < best[0]:
            best = (tot, t)
    e.set_timing(False)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(e.stream)
    e.clear()
    ev0.record(st)
    e.process_device(x.data_ptr(), n, y.data_ptr(), n, n, sync=False)
    ev1.record(st)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if ref is None:
        ref = y.clone()
        err = 0.0
    else:
        err = float((y - ref).abs().max() / ref.abs().max())
    print(f"block {B:5d}: P = {e.stages()[0]['partitions']:4d}  step {ms:8.3f} ms  ({n / ms / 1e3:9.1f} M stereo frames/s)  "
          f"sweep {best[1]['cmac_ms']:.3f}  fwd {best[1]['fft_ms']:.3f}  inv {best[1]['ifft_ms']:.3f}  max diff vs block 512: {err:.2e}", flush=True)
    e.close()
    del y
    torch.cuda.empty_cache()
