#!/bin/bash
# Multi-GPU bench (charged N x wall time!).  Usage: gpurun --gpus N --timeout 600 -- 'bash scripts/gpu_r02_multi.sh N [extra bench args]'
set -x
N=${1:-2}; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_g$N.txt 2>&1
df -h /dev/shm | tail -1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/r02_bench_g$N.json 2> gpurun_out/r02_bench_g$N.err
echo "rc=$?"; cut -c1-400 gpurun_out/r02_bench_g$N.json; grep -v "^$" gpurun_out/r02_bench_g$N.err | grep -vi "OMP_NUM\|\*\*\*\*" | tail -12
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench_g$N.json').read().splitlines() if l.startswith('{')][-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], '\nparity', d['parity'], '\nir120', {k:d['ir120'].get(k) for k in ('value','ms_per_step','parity','fp32_frac','step_share')}, '\nshare', d['roofline']['step_share'], d['roofline']['frac'])
except Exception as ex:
    print('no json', ex)
PY
