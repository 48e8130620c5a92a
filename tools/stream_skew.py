"""One measurement of the streaming sweep on config 5 (120 s IR): variant 103 (equal slices) and 108 (skewed slices,
B200CONV_STREAM_SKEW percent from the environment).  Run once per skew value (the value is read once per process)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reevr_b200.convolver import Engine  # noqa: E402
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402

C, B, L = 2, 512, 120 * 48000
irs = [synth_ir(L, c) for c in range(C)]
x = torch.from_numpy(np.stack([synth_input(B * 80, c) for c in range(C)])).cuda()
y = torch.empty((C, B), device="cuda")
for v in (103, 108):
    e = Engine(C, cmac_variant=v)
    e.init_uniform(B, irs)
    P = e.stages()[0]["partitions"]
    for i in range(8):
        e.process_device(x[:, i * B:].data_ptr(), x.shape[1], y.data_ptr(), B, B, sync=True)
    e.set_timing(True)
    ts = []
    for i in range(8, 72):
        e.process_device(x[:, i * B:].data_ptr(), x.shape[1], y.data_ptr(), B, B, sync=True)
        ts.append(e.last_timing()["cmac_ms"])
    alg = 16 * P * (B + 1) * C
    print(f"skew {os.environ.get('B200CONV_STREAM_SKEW', 'default')}% v{v}: median {np.median(ts) * 1e3:.2f} us  mean {np.mean(ts) * 1e3:.2f} us  "
          f"min {min(ts) * 1e3:.2f} us  -> {alg / (np.mean(ts) * 1e-3) / 1e9:.0f} GB/s (mean)", flush=True)
    e.close()
