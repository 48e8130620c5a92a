set -x
mkdir -p gpurun_out
cap() {  # cap <name> <kernel regex> <skip> <cmd...>
  name=$1; shift; rx=$1; shift; sk=$1; shift
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $sk -c 1 -f -o /tmp/$name "$@" > gpurun_out/${name}_run.log 2>&1
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/$name.txt
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-300 gpurun_out/bench_final.json; tail -3 gpurun_out/bench_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-stream > gpurun_out/ncu_launches_final.log 2>&1
cap prof_final_cmac k_cmac_batch2 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_final_fwd k_fwd_fft 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_final_inv k_inv_fft 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-stream
cap prof_final_stream_cfg5 k_cmac_stream_rows 280 python tools/stream_bench.py
timeout 200 python tools/stream_bench.py > gpurun_out/stream_final.txt 2>&1
du -sh gpurun_out; ls -la gpurun_out | tail -16
