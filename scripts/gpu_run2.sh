set -x
mkdir -p gpurun_out
./tools/ubench_fma > gpurun_out/ubench_fma.txt 2>&1; cat gpurun_out/ubench_fma.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -8 gpurun_out/pytest_gpu2.log
timeout 900 python tools/sweep.py > gpurun_out/sweep2.txt 2>&1; cat gpurun_out/sweep2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fft -s 3 -c 1 -f -o gpurun_out/prof_fwdfft_r01 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --blocks 4736 > gpurun_out/ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_inv_fft -s 3 -c 1 -f -o gpurun_out/prof_invfft_r01 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --blocks 4736 > gpurun_out/ncu_inv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cmac_batch2 -s 3 -c 1 -f -o gpurun_out/prof_cmac2_v22_r01 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --variant 22 > gpurun_out/ncu_c22.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cmac_batch2 -s 3 -c 1 -f -o gpurun_out/prof_cmac2_v21_r01 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --variant 21 > gpurun_out/ncu_c21.log 2>&1
ls -la gpurun_out
