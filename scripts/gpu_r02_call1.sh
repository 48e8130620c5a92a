#!/bin/bash
# round-2 call 1: run everything that was written blind in round 1 + sanitizer evidence + a baseline bench
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r02_smi.txt
nproc >> gpurun_out/r02_smi.txt; lscpu | head -30 >> gpurun_out/r02_smi.txt; numactl -H >> gpurun_out/r02_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_1.log
tail -5 gpurun_out/r02_pytest_gpu_1.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity.py tests/test_distributed.py -m gpu -q -k "golden or mixdown or hot_swap or slot_exchange" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck.log
tail -4 gpurun_out/r02_sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -q -k "golden" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck.log
tail -4 gpurun_out/r02_sanitizer_racecheck.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_base.json 2> gpurun_out/r02_bench_n1_base.err; cut -c1-400 gpurun_out/r02_bench_n1_base.json; tail -3 gpurun_out/r02_bench_n1_base.err
