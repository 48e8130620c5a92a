set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -6 gpurun_out/pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-900 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cmac_batch2 -s 3 -c 1 -f -o gpurun_out/prof_final_cmac python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fwd_fft -s 3 -c 1 -f -o gpurun_out/prof_final_fwd python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_fft -s 3 -c 1 -f -o gpurun_out/prof_final_inv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cmac_stream_rows -s 280 -c 1 -f -o gpurun_out/prof_final_stream_cfg5 python tools/stream_bench.py > gpurun_out/stream_final.txt 2>&1
ls -la gpurun_out | tail -8
