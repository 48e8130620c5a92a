#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_7.log
tail -6 gpurun_out/r02_pytest_gpu_7.log
timeout 300 python tools/sweep.py --blocks 28152 --variants 22 33 34 21 23 28 > gpurun_out/r02_sweep_variants.txt 2>&1; cat gpurun_out/r02_sweep_variants.txt
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-traffic > gpurun_out/r02_bench_n1_e.json 2> gpurun_out/r02_bench_n1_e.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_e.json')); print(d['value'], d['ms_per_step'], d['roofline']['step_share'], d['roofline']['frac'], d['e2e']['value'], d['roofline_stream'].get('frac'), d['realtime_process'])"
