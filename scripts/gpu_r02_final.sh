#!/bin/bash
# round-2 evidence run on one GPU (~10 GPU-minutes): full GPU test pass, sanitizer on the new kernels, the bench line,
# the reference arm, ncu launch list + full captures of the dominant kernels.  Usage: gpurun --timeout 3000 -- 'bash scripts/gpu_r02_final.sh'
set -x
mkdir -p gpurun_out
cap() {  # cap <name> <kernel regex> <skip> <cmd...>
  name=$1; shift; rx=$1; shift; sk=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.txt
  rm -f /tmp/$name.ncu-rep gpurun_out/${name}_run.log
}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_realtime.py tests/test_fft512.py tests/test_sliced.py tests/test_chain.py tests/test_parity.py -m gpu -q -k "realtime or callback or cluster or split or mixdown or fft512 or uniform_512 or impulse or two_stage_with or sliced or chain_against or streaming" > gpurun_out/r02_sanitizer_memcheck2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck2.log
tail -4 gpurun_out/r02_sanitizer_memcheck2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_fft512.py tests/test_realtime.py -m gpu -q -k "uniform_512 or two_stage_callback or quad" > gpurun_out/r02_sanitizer_racecheck2.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck2.log
tail -4 gpurun_out/r02_sanitizer_racecheck2.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2>&1; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-stream --no-traffic --no-ir120 --no-parity > gpurun_out/ncu_launches.log 2>&1
BARGS="--steps 1 --warmup 3 --no-cpu --no-e2e --no-stream --no-traffic --no-ir120 --no-parity"
cap r02_prof_cmac k_cmac_batch2 3 python bench.py $BARGS
cap r02_prof_fwd512b k_fwd_fft512 3 python bench.py $BARGS
cap r02_prof_inv512b k_inv_fft512 3 python bench.py $BARGS
cap r02_prof_stream_tma k_cmac_stream_tma 9 python bench.py --probe stream
cap r02_prof_rt k_rt_block 150 python tools/stream_bench.py
timeout 300 python tools/stream_bench.py > gpurun_out/r02_stream_bench.txt 2>&1; tail -12 gpurun_out/r02_stream_bench.txt | cut -c1-200
du -sh gpurun_out
