#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_6.log
tail -5 gpurun_out/r02_pytest_gpu_6.log
timeout 300 python tools/stream_bench.py > gpurun_out/r02_stream_bench2.txt 2>&1; tail -12 gpurun_out/r02_stream_bench2.txt | cut -c1-220
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-traffic --no-ir120 > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_d.json')); print(d['value'], d['roofline']['step_share'], d['roofline_stream'].get('frac'), d['realtime_process'])"
