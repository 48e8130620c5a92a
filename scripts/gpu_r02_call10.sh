#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_realtime.py tests/test_cpp_dropin.py tests/test_chain.py -m gpu -q > gpurun_out/r02_pytest_gpu_10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_10.log
tail -4 gpurun_out/r02_pytest_gpu_10.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fft512.py tests/test_realtime.py -m gpu -q -k "uniform_512 or two_stage_callback or quad or split" > gpurun_out/r02_sanitizer_racecheck3.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck3.log
tail -4 gpurun_out/r02_sanitizer_racecheck3.log
