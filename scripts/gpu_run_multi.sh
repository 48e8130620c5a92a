set -x
N=${1:-2}
EXTRA=${2:-}
mkdir -p gpurun_out
if [ "$EXTRA" = "test" ]; then timeout 300 python -m pytest tests/test_distributed.py -m gpu -q -x 2>&1 | tail -5; fi
NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --also-ir120 > gpurun_out/bench_p2p_g$N.json 2> gpurun_out/bench_p2p_g$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/bench_p2p_g$N.json; grep -v "^$" gpurun_out/bench_p2p_g$N.err | grep -vi "OMP_NUM\|\*\*\*\*" | tail -12
