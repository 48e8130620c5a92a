"""Times the streaming FDL sweep variants (one block per launch) on one GPU: 101 = register batches
(k_cmac_stream_rows), 102-105 = TMA ring (k_cmac_stream_tma: stages x CTAs/SM = 4x3, 6x2, 12x1, 2x6)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reevr_b200.convolver import Engine  # noqa: E402
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402

pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
shapes = [("cfg5_120s_B512", 2, 512, 120 * 48000), ("metric_10s_B512", 2, 512, 480000), ("B128_30s", 2, 128, 30 * 48000),
          ("B8192_tail_60s", 2, 8192, 60 * 48000)]
for name, C, B, L in shapes:
    irs = [synth_ir(L, c) for c in range(C)]
    x = torch.from_numpy(np.stack([synth_input(B * 64, c) for c in range(C)])).cuda()
    y = torch.empty((C, B), device="cuda")
    ref = None
    for v in (101, 103, 104, 106, 107):
        e = Engine(C, cmac_variant=v)
        e.init_uniform(B, irs)
        P = e.stages()[0]["partitions"]
        for i in range(8):
            e.process_device(x[:, i * B:].data_ptr(), x.shape[1], y.data_ptr(), B, B, sync=True)
        e.set_timing(True)
        ts = []
        for i in range(8, 56):
            e.process_device(x[:, i * B:].data_ptr(), x.shape[1], y.data_ptr(), B, B, sync=True)
            ts.append(e.last_timing()["cmac_ms"])
        out = y.clone()
        if ref is None:
            ref = out
        err = float((out - ref).abs().max() / ref.abs().max())
        cm = float(np.median(ts))
        alg = 16 * P * (B + 1) * C
        print(f"{name} P={P} v{v}: {cm * 1e3:7.2f} us  {alg / (cm * 1e-3) / 1e9:7.1f} GB/s  frac {alg / (cm * 1e-3) / 1e9 / PEAK:.3f}  "
              f"min {min(ts) * 1e3:.2f} us  err_vs_101 {err:.1e}", flush=True)
        e.close()
