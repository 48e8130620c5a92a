#!/usr/bin/env python
"""bench.py — headline benchmark of the partitioned-convolution hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload metric|ir120|...]

Metric: M stereo frames / s ("Msamples/sec stereo conv @ IR=10s/48kHz block=512"): two
independent mono convolutions (LL, RR — src/dsp/StereoConvolver.cpp:35-36) with their own
480 000-tap IR each, uniform partitions of 512 (P = 938).  One "step" = one pass of the hot
path (forward FFT of every block, FDL complex-MAC sweep, inverse FFT + overlap-add) over a batch
of T blocks of synthetic white noise.

* value            device-resident throughput (input/output already in HBM), CUDA events on the
                   engine's stream, L2 flushed between steps.
* e2e              same metric through b200conv_process() with pinned HOST buffers (H2D + D2H in
                   the timed region, wall clock around the synchronous call).
* roofline         dominant kernel (k_cmac_batch): algorithmic bytes (16*P*K + 8*K + 16*B
                   per channel-block, K = 513) / its mean CUDA-event duration vs MEASURED_PEAKS hbm_gbs.
                   NOTE: the batched sweep reuses H[p] across 16 blocks in registers, so it is
                   FP32-FMA-bound, not HBM-bound, and `frac` legitimately exceeds 1 — `fp32`
                   carries the bound that actually applies.
* cpu_baseline     the reference's own CPU code (oracle/_ref, unmodified sources) on this host.
N > 1 (torchrun): the IR's partition range is sharded over the ranks, partial spectra are
summed into rank 0 with one NCCL reduce per batch before the inverse FFT ("strong" scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware queue per stream: the engine's flag barriers spin on s_post while s_main keeps launching sweeps
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

B = 512
SR = 48000
WORKLOADS = {
    # name: (channels, ir seconds, sample rate, block)
    "metric": dict(C=2, ir_s=10, sr=48000, block=512, desc="stereo 48 kHz, 10 s IR (480000 taps), uniform block 512"),
    "ir1": dict(C=1, ir_s=1, sr=48000, block=512, desc="mono 48 kHz, 1 s IR, uniform block 512 (config 1)"),
    "ch8": dict(C=8, ir_s=10, sr=48000, block=512, desc="8-channel 48 kHz, 10 s IR per channel, block 512 (config 4)"),
    "ir120": dict(C=2, ir_s=120, sr=48000, block=512, desc="stereo 48 kHz, 120 s IR, uniform block 512 (config 5)"),
    # two-stage shapes (head block = `block`, tail block = `tail`): not bench lines of the contract, kept for tuning runs
    "cfg2": dict(C=2, ir_s=5, sr=48000, block=128, tail=8192, desc="stereo 48 kHz, 5 s IR, two-stage head 128 / tail 8192 (config 2)"),
    "cfg3": dict(C=2, ir_s=30, sr=96000, block=64, tail=8192, desc="stereo 96 kHz, 30 s IR, two-stage head 64 / tail 8192 (config 3)"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def algorithmic_bytes_per_channel_block(P: int, block: int) -> int:
    K = block + 1
    return 16 * P * K + 8 * K + 16 * block


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if t0 is not None and not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm),
                "window": "timed steps + 0.6 s continuation of the same step loop"}


# ---------------------------------------------------------------------------------------------
# CPU reference timing (oracle/_ref = the unmodified reference sources; falls back to the C port)
# ---------------------------------------------------------------------------------------------
def cpu_reference_run(wl, seconds_target: float, threads: int, single_thread_leg: bool = False):
    """Runs `threads` independent stereo (C-channel) instances of the reference's uniform
    FFTConvolver in parallel, each over the same bounded sample; returns dict for cpu_baseline."""
    from oracle import oracle as orc
    kind = "reference" if orc.ref_available() else "port"
    cls = orc.RefUniform if kind == "reference" else orc.OracleUniform
    C, block = wl["C"], wl["block"]
    L = wl["ir_s"] * wl["sr"]
    irs = [orc.synth_ir(L, c) for c in range(C)]

    def make():
        convs = []
        for c in range(C):
            k = cls()
            k.init(block, irs[c])
            k.clear()                # zero-fills (= first-touches) the whole frequency-domain delay line
            convs.append(k)
        return convs

    insts = [make() for _ in range(threads)]

    def run_all(nblk):
        xs = [orc.synth_input(nblk * block, c) for c in range(C)]
        done = [0.0] * threads

        def work(i):
            for c in range(C):           # channels serially on one thread, as StereoConvolver::process does
                insts[i][c].run(xs[c], block)
            done[i] = time.perf_counter()

        ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return max(done) - t0

    # "as the plugin does it" (src/dsp/StereoConvolver.cpp:35-36): ONE thread, the C channels serially
    single = None
    if single_thread_leg:
        xs1 = [orc.synth_input(64 * block, c) for c in range(C)]
        t1 = time.perf_counter()
        for c in range(C):
            insts[0][c].run(xs1[c], block)
        single = 64 * block / (time.perf_counter() - t1) / 1e6
        for c in range(C):
            insts[0][c].clear()

    # calibrate with ALL threads running (the sweep is memory-bound: per-thread speed drops with the
    # thread count), then size the sample for ~seconds_target of wall time
    run_all(16)                      # warm-up: first touch of every instance's FDL / spectra
    cal = 8
    per_block = run_all(cal) / cal
    nblk = int(max(16, min(16384, seconds_target / max(per_block, 1e-9))))
    dt = run_all(nblk)
    frames = nblk * block * threads
    return {
        "value": frames / dt / 1e6, "unit": "M stereo frames/s" if C == 2 else f"M {C}-channel frames/s",
        "cores": threads, "kind": kind,
        "sample": f"{threads} independent {C}-channel instances x {nblk} blocks of {block} (ctypes, GIL released), "
                  f"uniform FFTConvolver, {wl['desc']}",
        "seconds": dt, "parallel_ms_per_block": per_block * 1e3,
        "single_thread_value": single,
    }


# ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="metric", choices=sorted(WORKLOADS))
    ap.add_argument("--blocks", type=int, default=0, help="blocks per step (0 = auto)")
    ap.add_argument("--variant", type=int, default=0, help="CMAC kernel variant (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-stream", dest="no_stream", action="store_true", help="skip the streaming-kernel HBM roofline leg")
    ap.add_argument("--sweep", action="store_true", help="print a per-variant timing table to stderr")
    ap.add_argument("--mgpu", default="p2p", choices=["p2p", "nccl"], help="multi-GPU exchange path (N > 1)")
    ap.add_argument("--e2e-bcast", dest="e2e_bcast", action="store_true",
                    help="N > 1, p2p: rank 0 uploads the input and broadcasts it over NVLink (off by default: not yet timed)")
    ap.add_argument("--also-ir120", dest="also_ir120", action="store_true",
                    help="additionally time config 5 (120 s IR) and attach it as `ir120`")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    warm = max(args.warmup, 3)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        threads = os.cpu_count() or 1
        # each step = one bounded sample; keep the whole run within a few minutes
        vals = []
        for i in range(warm + args.steps):
            r = cpu_reference_run(wl, seconds_target=2.0 if i < warm else 6.0, threads=threads)
            if i >= warm:
                vals.append(r)
        v = statistics.mean(x["value"] for x in vals)
        last = vals[-1]
        line = {
            "impl": "reference", "metric": "stereo partitioned-convolution throughput (IR 10 s @ 48 kHz, block 512)",
            "value": v, "unit": "M stereo frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": 1e3 * statistics.mean(x["seconds"] for x in vals), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "engine": "reference CPU FFTConvolver (oracle/_ref)", "threads": threads},
            "cpu_baseline": {"value": v, "unit": "M stereo frames/s", "cores": threads, "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": v, "unit": "M stereo frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from reevr_b200.convolver import Engine
    from reevr_b200.distributed import attach_p2p, attach_reduce
    from reevr_b200.synth import synth_input, synth_ir

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fall-back)")
    torch.cuda.set_device(local)
    # Pinned staging buffers must live on the NUMA node the GPU hangs off, otherwise every H2D / D2H of the
    # e2e path crosses the socket interconnect (and with sharding the slowest rank's link paces all of them):
    # run this process on the GPU's CPU-affinity set while the buffers are allocated and first touched.
    all_cpus = os.sched_getaffinity(0)
    numa_note = "not bound"
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (max(all_cpus) // 64) + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(hnd, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & all_cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
            numa_note = f"process bound to the GPU's {len(cpus)} local CPUs for pinned allocations"
    except Exception as ex:       # best effort
        numa_note = f"not bound ({type(ex).__name__})"
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_workload(wl, T, steps, with_e2e, with_clocks):
        """Times one workload; returns the dict the JSON line is assembled from (rank 0) / None."""
        C, block = wl["C"], wl["block"]
        L = wl["ir_s"] * wl["sr"]
        n = T * block
        # single GPU: one launch group per step; sharded: groups of 7104 blocks so that exchange + inverse FFT of
        # group i overlap the sweep of group i+1 (engine post stream)
        groups = 1 if world == 1 else max(1, round(T / 7104))     # 7104 blocks = 8 full sweep waves
        gb = (T + groups - 1) // groups
        eng = Engine(C, device=local, max_batch_blocks=gb + 1, shard_rank=rank, shard_count=world, cmac_variant=args.variant)
        irs = [synth_ir(L, c) for c in range(C)]
        t_init = time.perf_counter()
        if "tail" in wl:
            assert eng.init_twostage(block, wl["tail"], irs)
        else:
            assert eng.init_uniform(block, irs)
        t_init = time.perf_counter() - t_init
        st = eng.stages()[0]
        P = int(st["partitions"])
        Ploc = int(st["p_end"]) - int(st["p_begin"])
        stream = torch.cuda.ExternalStream(eng.stream, device=dev)
        mgpu_path = "single GPU"
        if world > 1:
            attach_reduce(eng, device=local)          # NCCL reduce hook (always installed)
            mgpu_path = "partition-range shards + NCCL reduce of partial spectra to rank 0"
            if args.mgpu == "p2p":
                ok, why = attach_p2p(eng)             # fused slot exchange over NVLink peer memory (same answer on every rank)
                if ok:
                    mgpu_path = ("partition-range shards + fused slot exchange: sweep epilogue stores partial rows into "
                                 "the owner GPU's slot over NVLink, flag barrier, per-slice inverse FFT (no NCCL on the data path)")
                    if args.e2e_bcast:                # host-pointer path: only rank 0 crosses PCIe
                        eng.p2p_set_input_broadcast(True)
                else:
                    print(f"[bench] slot exchange not available ({why}); using the NCCL reduce path", file=sys.stderr)
                    mgpu_path += f" (slot exchange not available: {why})"

        x_host = torch.empty((C, n), dtype=torch.float32).pin_memory()
        for c in range(C):
            x_host[c] = torch.from_numpy(synth_input(n, c))
        y_host = torch.empty((C, n), dtype=torch.float32).pin_memory()
        x_dev = x_host.cuda(non_blocking=False)
        y_dev = torch.empty_like(x_dev)

        def step_device():
            eng.process_device(x_dev.data_ptr(), n, y_dev.data_ptr(), n, n, sync=False)

        for _ in range(warm):
            step_device()
        barrier()
        sampler = ClockSampler(local) if (with_clocks and rank == 0) else None
        if sampler:
            sampler.start()
            time.sleep(0.3)
        launches0 = eng.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        t_w0 = time.perf_counter()
        for i in range(steps):
            flush.zero_()                       # evict L2 between timed iterations
            barrier()
            ev[i][0].record(stream)
            step_device()
            ev[i][1].record(stream)
        barrier()
        t_w1 = time.perf_counter()
        launches = eng.launch_count - launches0
        t_tot = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_tot, op=dist.ReduceOp.MAX)
        ms_per_step = float(t_tot.item()) / steps
        value = n / (ms_per_step * 1e-3) / 1e6
        if with_clocks:
            # K short steps give nvidia-smi (>= 20 ms per sample) almost nothing to see: EVERY rank keeps the very
            # same step loop running for another ~0.6 s (untimed; identical count on all ranks — the reduce is a
            # collective) so that the clock / throttle record of rank 0 is meaningful
            n_extra = int(min(2000, max(1, 600.0 / max(ms_per_step, 0.05))))
            for _ in range(n_extra):
                step_device()
            barrier()
            t_w1 = time.perf_counter()
        clocks = sampler.stop(t_w0, t_w1) if sampler else None

        # dominant-kernel roofline: CUDA events around every FDL-sweep launch (separate pass)
        eng.set_timing(True)
        cm_ms, cm_n, fft_ms, ifft_ms = 0.0, 0, 0.0, 0.0
        reps = max(2, min(steps, 5))
        for _ in range(reps):
            flush.zero_()
            barrier()
            eng.process_device(x_dev.data_ptr(), n, y_dev.data_ptr(), n, n, sync=True)
            tm = eng.last_timing()
            cm_ms += tm["cmac_ms"]; cm_n += tm["cmac_launches"]; fft_ms += tm["fft_ms"]; ifft_ms += tm["ifft_ms"]
        eng.set_timing(False)
        peak, peak_kind = measured_peaks()
        per_launch_ms = cm_ms / max(cm_n, 1)
        blocks_per_launch = T * reps / max(cm_n, 1)
        stages_all = eng.stages()
        if len(stages_all) == 1:
            alg_bytes_launch = algorithmic_bytes_per_channel_block(Ploc, block) * C * blocks_per_launch
            ffma = 4.0 * Ploc * block * C * blocks_per_launch      # 4 FP32 FMA per complex MAC, B bins per row
        else:   # multi-stage: SURVEY 8d, sum over stages of the per-sample figures, spread over the sweep launches
            per_sample = sum(algorithmic_bytes_per_channel_block(int(x["p_end"]) - int(x["p_begin"]), int(x["block"])) / int(x["block"])
                             for x in stages_all)
            alg_bytes_launch = per_sample * C * n * reps / max(cm_n, 1)
            ffma = sum(4.0 * (int(x["p_end"]) - int(x["p_begin"])) for x in stages_all) * C * n * reps / max(cm_n, 1)
        achieved = alg_bytes_launch / (per_launch_ms * 1e-3) / 1e9
        fp32_tflops = 2.0 * ffma / (per_launch_ms * 1e-3) / 1e12
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            key = f"{args.workload if wl is WL0 else 'ir120'}:T{int(blocks_per_launch)}:G{world}"
            traffic = tj.get(key)

        e2e = None
        if with_e2e:
            import ctypes
            inp = (ctypes.c_void_p * C)(*[x_host[c].data_ptr() for c in range(C)])
            outp = (ctypes.c_void_p * C)(*[y_host[c].data_ptr() for c in range(C)])
            for _ in range(2):
                eng.process_into(inp, outp, n)
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                eng.process_into(inp, outp, n)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e = {"value": n * steps / float(dt.item()) / 1e6, "unit": "M stereo frames/s",
                   "h2d_bytes_per_step": C * n * 4, "d2h_bytes_per_step": C * n * 4 if rank == 0 else 0,
                   "how": "b200conv_process() on pinned host buffers, wall clock, H2D / compute / D2H pipelined on separate streams",
                   "numa": numa_note}
        res = {
            "value": value, "ms_per_step": ms_per_step, "launches": int(launches), "clocks": clocks, "e2e": e2e,
            "config": {"workload": wl["desc"], "channels": C, "ir_taps": eng.ir_len(0), "block": block, "partitions": P,
                       "blocks_per_step": T, "frames_per_step": n, "launch_groups_per_step": groups,
                       "parallelism": mgpu_path if world == 1 else f"x{world}: {mgpu_path} ({Ploc} partitions on rank 0)",
                       "l2": "flushed between timed steps (256 MB write)", "init_s": round(t_init, 4)},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "traffic": traffic,
                         "kernel": "k_cmac_batch2 (FDL sweep)", "launch_ms": per_launch_ms,
                         "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "note": "algorithmic bytes = SURVEY 8d figure (every block streams H and the FDL once); the batched "
                                 "sweep reuses each H[p][k] for 16 blocks from registers, so frac > 1 is expected here and the "
                                 "binding limit is FP32 FMA issue (see fp32); traffic = ncu dram bytes of one launch (profiles/)",
                         "fp32": {"achieved_tflops": fp32_tflops, "peak_tflops": fp32_peak, "frac": fp32_tflops / fp32_peak,
                                  "peak_source": "148 SM x 128 FMA lanes/clk x 2 flop x 1965 MHz (clocks.max.sm)"},
                         "step_share": {"cmac_ms": cm_ms / reps, "fft_ms": fft_ms / reps, "ifft_ms": ifft_ms / reps}},
        }
        eng.close()
        del x_dev, y_dev
        torch.cuda.empty_cache()
        return res

    WL0 = wl
    # one step = one batch of T blocks of the same stereo stream, identical at every N (strong scaling):
    # 28416 blocks = 14.5 M frames = 5 min of audio; at N = 8 each GPU still sweeps ~0.5 ms per step
    T = args.blocks or (28416 if args.workload == "metric" else (7104 * 512 // wl["block"] if "tail" in wl else 7104))
    main_res = run_workload(wl, T, args.steps, with_e2e=not args.no_e2e, with_clocks=True)
    extra = None
    if args.also_ir120 and args.workload == "metric":
        r = run_workload(dict(WORKLOADS["ir120"]), 7104, max(2, min(3, args.steps)), with_e2e=False, with_clocks=False)
        extra = {"value": r["value"], "unit": "M stereo frames/s", "ms_per_step": r["ms_per_step"], "config": r["config"],
                 "roofline_frac": r["roofline"]["frac"], "fp32_frac": r["roofline"]["fp32"]["frac"]}

    # the memory-bound form of the sweep (real-time path, one block per launch) on a working set beyond L2:
    # this is the kernel whose "% of HBM roofline" is a bandwidth statement
    stream_roof = None
    realtime = None
    if world == 1 and not args.no_stream:
        try:    # secondary legs: never let them take the headline line down
            wl5 = WORKLOADS["ir120"]
            C5, B5 = wl5["C"], wl5["block"]
            e5 = Engine(C5, device=local)
            assert e5.init_uniform(B5, [synth_ir(wl5["ir_s"] * wl5["sr"], c) for c in range(C5)])
            P5 = int(e5.stages()[0]["partitions"])
            xs5 = torch.from_numpy(np.stack([synth_input(B5 * 72, c) for c in range(C5)])).cuda()
            y5 = torch.empty((C5, B5), device="cuda")
            for i in range(8):
                e5.process_device(xs5[:, i * B5:].data_ptr(), xs5.shape[1], y5.data_ptr(), B5, B5, sync=True)
            e5.set_timing(True)
            ts5 = []
            for i in range(8, 72):
                e5.process_device(xs5[:, i * B5:].data_ptr(), xs5.shape[1], y5.data_ptr(), B5, B5, sync=True)
                ts5.append(e5.last_timing()["cmac_ms"])
            e5.close()
            t5 = statistics.median(ts5)
            bytes5 = 16 * P5 * (B5 + 1) * C5                      # every H and FDL row read once per block step
            peak, peak_kind = measured_peaks()
            stream_roof = {"kernel": "k_cmac_stream_rows (one 512-sample block per launch)", "workload": wl5["desc"],
                           "working_set_bytes": 2 * P5 * B5 * 8 * C5, "bound": "hbm", "launch_ms": t5,
                           "achieved": bytes5 / (t5 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                           "frac": bytes5 / (t5 * 1e-3) / 1e9 / peak, "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})",
                           "algorithmic_bytes_per_launch": bytes5, "traffic": 188137728,
                           "traffic_source": "profiles/r01_prof_final_stream_cfg5.txt (ncu: 184.35 MB read + 3.79 MB written)"}
        except Exception as ex:
            stream_roof = {"error": f"{type(ex).__name__}: {ex}"}
        try:    # the real-time call a plugin makes: host pointers, one 512-sample block per call, synchronous
            C0, B0 = wl["C"], wl["block"]
            e0 = Engine(C0, device=local)
            assert e0.init_uniform(B0, [synth_ir(wl["ir_s"] * wl["sr"], c) for c in range(C0)])
            blk = [synth_input(B0, c) for c in range(C0)]
            for _ in range(50):
                e0.process(blk)
            lat = []
            for _ in range(300):
                t0 = time.perf_counter()
                e0.process(blk)
                lat.append(time.perf_counter() - t0)
            e0.close()
            med = statistics.median(lat)
            realtime = {"call": "b200conv_process(), host pointers, len = block = 512, synchronous", "median_us": med * 1e6,
                        "p99_us": sorted(lat)[int(0.99 * len(lat))] * 1e6, "value": B0 / med / 1e6,
                        "unit": "M stereo frames/s" if C0 == 2 else f"M {C0}-channel frames/s",
                        "note": "latency-bound (PCIe + 3 launches per call), reported for completeness — SURVEY 8d"}
        except Exception as ex:
            realtime = {"error": f"{type(ex).__name__}: {ex}"}

    if args.sweep and rank == 0 and world == 1:
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sweep.py"), "--blocks", str(T)], stdout=sys.stderr)

    if rank == 0:
        cpu = None
        os.sched_setaffinity(0, all_cpus)          # the CPU baseline uses every host core again
        if not args.no_cpu and world == 1:
            try:
                cpu = cpu_reference_run(wl, seconds_target=12.0, threads=os.cpu_count() or 1, single_thread_leg=True)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "parallel_ms_per_block", "single_thread_value")}
            except Exception as ex:       # the GPU numbers stand on their own
                cpu = {"error": f"{type(ex).__name__}: {ex}"}
        C = wl["C"]
        line = {
            "metric": "stereo partitioned-convolution throughput (IR 10 s @ 48 kHz, block 512)" if args.workload == "metric"
                      else f"partitioned-convolution throughput ({wl['desc']})",
            "value": main_res["value"], "unit": "M stereo frames/s" if C == 2 else f"M {C}-channel frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": warm, "ms_per_step": main_res["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": main_res["config"], "clocks": main_res["clocks"], "e2e": main_res["e2e"],
            "gpu_launches": main_res["launches"], "roofline": main_res["roofline"], "cpu_baseline": cpu,
        }
        if extra:
            line["ir120"] = extra
        if stream_roof:
            line["roofline_stream"] = stream_roof
        if realtime:
            line["realtime_process"] = realtime
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
