set -x
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_fma tools/ubench_fma.cu && /tmp/ubench_fma > gpurun_out/ubench_fma.txt 2>&1; cat gpurun_out/ubench_fma.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
tail -12 gpurun_out/pytest_gpu3.log
timeout 600 python tools/sweep.py --variants 12 21 22 28 > gpurun_out/sweep3.txt 2>&1; cat gpurun_out/sweep3.txt
timeout 600 python tools/stream_bench.py > gpurun_out/stream3.txt 2>&1; tail -30 gpurun_out/stream3.txt
timeout 900 python bench.py --steps 5 --warmup 3 --variant 22 --also-ir120 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; cat gpurun_out/bench3.json; tail -5 gpurun_out/bench3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench3_ref.json 2>&1; cat gpurun_out/bench3_ref.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fft -s 3 -c 1 -f -o gpurun_out/prof_fwdfft_r01b python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --blocks 4736 > gpurun_out/ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv_fft -s 3 -c 1 -f -o gpurun_out/prof_invfft_r01b python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --blocks 4736 > gpurun_out/ncu_inv.log 2>&1
ls -la gpurun_out | tail -12
