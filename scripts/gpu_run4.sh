set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu4.log
tail -6 gpurun_out/pytest_gpu4.log
timeout 600 python tools/sweep.py --variants 22 28 > gpurun_out/sweep4.txt 2>&1; cat gpurun_out/sweep4.txt
timeout 600 python tools/stream_bench.py > gpurun_out/stream4.txt 2>&1; grep -v STREAM_BENCH_JSON gpurun_out/stream4.txt | tail -14
timeout 900 python bench.py --steps 5 --warmup 3 --also-ir120 --no-cpu > gpurun_out/bench4.json 2> gpurun_out/bench4.err; cat gpurun_out/bench4.json | cut -c1-1500; tail -5 gpurun_out/bench4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cmac_stream_rows -s 20 -c 1 -f -o gpurun_out/prof_stream_cfg5_r01 python tools/stream_bench.py > gpurun_out/ncu_stream.log 2>&1
ls -la gpurun_out | tail -5
