#!/bin/bash
set -x
mkdir -p gpurun_out
cap() {  # cap <name> <kernel regex> <skip> <cmd...>
  name=$1; shift; rx=$1; shift; sk=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.txt
  ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_4.log
tail -6 gpurun_out/r02_pytest_gpu_4.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu --no-traffic > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n1_b.json')); print(d['value'], d['ms_per_step'], d['roofline']['step_share'], d['e2e']['value'], d['roofline_stream'].get('frac'), d['realtime_process'])"
cap r02_prof_fwd512 k_fwd_fft512 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream --no-traffic --no-ir120 --no-parity
cap r02_prof_inv512 k_inv_fft512 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream --no-traffic --no-ir120 --no-parity
cap r02_prof_stream_tma k_cmac_stream_tma 9 python bench.py --probe stream
head -40 gpurun_out/r02_prof_inv512.txt
