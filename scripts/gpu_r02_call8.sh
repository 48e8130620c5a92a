#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py tests/test_chain.py tests/test_sliced.py -m gpu -q -x > gpurun_out/r02_pytest_gpu_8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_8.log
tail -4 gpurun_out/r02_pytest_gpu_8.log
timeout 600 python tools/stream_variants.py > gpurun_out/r02_stream_variants2.txt 2>&1; cat gpurun_out/r02_stream_variants2.txt | tail -24
