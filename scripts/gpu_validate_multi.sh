#!/bin/bash
# Multi-GPU bench (charged N x wall time!).  Usage: gpurun --gpus N --timeout 400 -- 'bash scripts/gpu_validate_multi.sh N [nccl|p2p]'
set -x
N=${1:-2}
PATHSEL=${2:-p2p}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_g$N.txt 2>&1
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 5 --warmup 3 --mgpu $PATHSEL --also-ir120 > gpurun_out/bench_${PATHSEL}_g$N.json 2> gpurun_out/bench_${PATHSEL}_g$N.err
echo "rc=$?"; cut -c1-300 gpurun_out/bench_${PATHSEL}_g$N.json; grep -v "^$" gpurun_out/bench_${PATHSEL}_g$N.err | grep -vi "OMP_NUM\|\*\*\*\*" | tail -8
