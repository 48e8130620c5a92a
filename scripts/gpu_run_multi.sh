set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_g$N.txt 2>&1
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --also-ir120 > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "rc=$?"; tail -c 2500 gpurun_out/bench_g$N.json; tail -15 gpurun_out/bench_g$N.err
