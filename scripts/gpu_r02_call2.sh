#!/bin/bash
# round-2 call 2: full GPU test pass + the restructured bench at N = 1 (parity objects, ir120 leg, traffic probe)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_2.log
tail -8 gpurun_out/r02_pytest_gpu_2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_n1_a.json; tail -5 gpurun_out/r02_bench_n1_a.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02_bench_ref_a.json 2>&1; cut -c1-300 gpurun_out/r02_bench_ref_a.json
