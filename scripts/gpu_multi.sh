#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus N`): the NCCL parity test and bench.py at N ranks.
# usage: scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; TAG=${2:-r01e}
mkdir -p gpurun_out
nvidia-smi -L
echo "== nccl test"; timeout 600 python -m pytest tests/test_sharding.py -m gpu -q -x 2>&1 | tail -4
echo "== bench N=$N"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
tail -c 3000 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err
free -g | head -2
