#!/bin/bash
# round-2 record run with the tensor-core sweep in the library (~5 GPU-minutes): full GPU test pass, the bench line,
# the reference arm, the ncu launch list of the bench command and one full capture of k_tc_sweep, memcheck of the
# standalone tensor-core harness, the shifted-window probe.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r02_tc_final.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -4 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["traffic"], d["roofline"]["step_share"], d["parity"]["ok"], d["parity"]["max_err_vs_ref"], d["ir120"]["value"], d["cpu_baseline"]["value"], d["roofline_stream"]["frac"], d["realtime_process"]["reevr_quad"]["median_us"], d["gpu_launches"], d["clocks"])
PY
timeout 200 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>&1; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_tc.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-stream --no-traffic --no-ir120 --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_tc_sweep -c 1 -f -o /tmp/tc_final tools/bin/tc_sweep_test 2 512 938 112608 > gpurun_out/tc_final_ncu.log 2>&1
python tools/ncu_summary.py /tmp/tc_final.ncu-rep gpurun_out/r02_prof_tc_sweep.txt
timeout 120 compute-sanitizer --tool memcheck --error-exitcode 9 tools/bin/tc_sweep_test 2 64 100 300 > gpurun_out/r02_sanitizer_tc_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_tc_memcheck.log
tail -4 gpurun_out/r02_sanitizer_tc_memcheck.log
timeout 60 tools/bin/tc_probe > gpurun_out/r02_tc_probe.txt 2>&1; tail -12 gpurun_out/r02_tc_probe.txt
du -sh gpurun_out
