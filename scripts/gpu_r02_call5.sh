#!/bin/bash
set -x
mkdir -p gpurun_out
# the one-launch real-time path first, on a short leash (cluster launch + DSMEM + zero-copy are new on real hardware)
timeout 300 python -m pytest tests/test_realtime.py -m gpu -q -x > gpurun_out/r02_pytest_rt.log 2>&1; rc=$?; echo "pytest rt rc=$rc" >> gpurun_out/r02_pytest_rt.log
tail -15 gpurun_out/r02_pytest_rt.log
if [ $rc -ne 0 ]; then
  echo "rt path failed: rerunning the suite with the path disabled"; export B200CONV_NO_RT=1
fi
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_5.log
tail -6 gpurun_out/r02_pytest_gpu_5.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-traffic > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_c.json')); print(d['value'], d['ms_per_step'], d['roofline']['step_share'], d['e2e']['value'], d['roofline_stream'].get('frac'), d['realtime_process'])"
timeout 300 python tools/stream_bench.py > gpurun_out/r02_stream_bench.txt 2>&1; tail -14 gpurun_out/r02_stream_bench.txt | cut -c1-300
