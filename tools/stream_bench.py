"""Streaming (real-time) path measurements on one GPU.

1. k_cmac_stream as an HBM/L2 stream: one block per launch, every H and FDL row read once —
   achieved GB/s = algorithmic bytes (16*P*K per channel-block) / CUDA-event time,
   for the metric shape (working set 15 MB, L2 resident) and config 5 (185 MB, beyond L2).
2. Host-pointer process() latency per call (H2D + 3 kernels + D2H + sync) for the shapes a
   REEV-R audio callback produces.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reevr_b200.convolver import Engine  # noqa: E402
from reevr_b200.synth import synth_input, synth_ir  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
out = {}

for name, secs in (("metric_10s", 10), ("cfg5_120s", 120)):
    C, B = 2, 512
    irs = [synth_ir(secs * 48000, c) for c in range(C)]
    e = Engine(C)
    e.init_uniform(B, irs)
    P = e.stages()[0]["partitions"]
    n = B
    x = torch.from_numpy(np.stack([synth_input(n * 64, c) for c in range(C)])).cuda()
    y = torch.empty((C, n), device="cuda")
    for i in range(8):
        e.process_device(x[:, i * n:].data_ptr(), x.shape[1], y.data_ptr(), n, n, sync=True)
    e.set_timing(True)
    ts = []
    for i in range(8, 56):
        e.process_device(x[:, i * n:].data_ptr(), x.shape[1], y.data_ptr(), n, n, sync=True)
        ts.append(e.last_timing())
    cm = np.median([t["cmac_ms"] for t in ts])
    ff = np.median([t["fft_ms"] for t in ts])
    iff = np.median([t["ifft_ms"] for t in ts])
    e.set_timing(False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 200
    for i in range(reps):
        e.process_device(x[:, (i % 60) * n:].data_ptr(), x.shape[1], y.data_ptr(), n, n, sync=False)
    torch.cuda.synchronize()
    per_call = (time.perf_counter() - t0) / reps
    alg = 16 * P * (B + 1) * C
    out[name] = dict(partitions=P, working_set_MB=2 * P * B * 8 * C / 1e6, cmac_us=cm * 1e3, fft_us=ff * 1e3, ifft_us=iff * 1e3,
                     sweep_GBps=alg / (cm * 1e-3) / 1e9, frac_of_hbm_peak=alg / (cm * 1e-3) / 1e9 / PEAK,
                     device_resident_block_step_us=per_call * 1e6,
                     stream_Mframes_s=n / per_call / 1e6)
    print(name, json.dumps(out[name]), flush=True)
    e.close()

# host-pointer latency, REEV-R shapes: two-stage head = host block, tail 8192, 10 s IR
lat = {}
for C in (1, 2, 4):
    for hb in (64, 128, 512):
        irs = [synth_ir(480000, c) for c in range(C)]
        e = Engine(C)
        e.init_twostage(hb, 8192, irs)
        xs = [synth_input(hb, c) for c in range(C)]
        for _ in range(300):          # crosses tail-block boundaries
            e.process(xs)
        ts = []
        for _ in range(400):
            t0 = time.perf_counter()
            e.process(xs)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e6
        lat[f"C{C}_head{hb}"] = dict(median_us=float(np.median(ts)), p99_us=float(np.percentile(ts, 99)), max_us=float(ts.max()),
                                     realtime_budget_us=hb / 48000 * 1e6)
        print(f"C{C}_head{hb}", json.dumps(lat[f"C{C}_head{hb}"]), flush=True)
        e.close()
out["process_latency"] = lat
print("STREAM_BENCH_JSON " + json.dumps(out))
