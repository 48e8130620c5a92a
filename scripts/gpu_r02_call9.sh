#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_realtime.py -m gpu -q -x > gpurun_out/r02_pytest_rt2.log 2>&1; echo "pytest rt rc=$?" >> gpurun_out/r02_pytest_rt2.log
tail -4 gpurun_out/r02_pytest_rt2.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_9.log
tail -5 gpurun_out/r02_pytest_gpu_9.log
timeout 300 python tools/stream_bench.py > gpurun_out/r02_stream_bench3.txt 2>&1; tail -12 gpurun_out/r02_stream_bench3.txt | cut -c1-260
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-traffic --no-ir120 > gpurun_out/r02_bench_n1_f.json 2> gpurun_out/r02_bench_n1_f.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_f.json')); print(d['value'], d['roofline']['step_share'], d['roofline_stream'].get('frac'), d['realtime_process'])"
