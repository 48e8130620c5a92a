set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu5.log
tail -12 gpurun_out/pytest_gpu5.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench5.json 2> gpurun_out/bench5.err; cut -c1-700 gpurun_out/bench5.json; tail -3 gpurun_out/bench5.err
