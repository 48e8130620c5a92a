set -x
N=${1:-4}
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_T28k_n$N.json 2> gpurun_out/bench_T28k_n$N.err
echo "rc=$?"; cut -c1-200 gpurun_out/bench_T28k_n$N.json; grep -v "^$" gpurun_out/bench_T28k_n$N.err | grep -vi "OMP_NUM\|\*\*\*\*" | tail -8
