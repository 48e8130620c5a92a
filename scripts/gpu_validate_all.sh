#!/bin/bash
# One-GPU validation + evidence run (≈ 6-8 GPU-minutes).  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_validate_all.sh'
# Writes only compact artefacts into gpurun_out/ (ncu reports are summarised and deleted: the merge-back limit is 64 MiB).
set -x
mkdir -p gpurun_out
cap() {  # cap <name> <kernel regex> <skip> <cmd...>
  name=$1; shift; rx=$1; shift; sk=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.txt
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
# memory checker on a small end-to-end case (uniform + two-stage + routing)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity.py -m gpu -q -k "golden or mixdown or hot_swap" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -3 gpurun_out/sanitizer_memcheck.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_reference.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-stream > gpurun_out/ncu_launches.log 2>&1
cap prof_cmac k_cmac_batch2 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_fwd k_fwd_fft 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_inv k_inv_fft 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_stream_cfg5 k_cmac_stream_rows 280 python tools/stream_bench.py
timeout 200 python tools/stream_bench.py > gpurun_out/stream.txt 2>&1
timeout 300 python tools/sweep.py --variants 12 21 22 28 > gpurun_out/sweep.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -20
