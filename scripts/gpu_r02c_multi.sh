#!/bin/bash
# round 2, call c: the shapes north_star quotes for 8 GPUs -- o=64,v=512 (configs[3], complete E(T) end to
# end) and o=100,v=800 (configs[4], hole-block mode, sampled) -- plus the default workload, N ranks.
N=${1:-8}; TAG=${2:-r02c}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
(nproc; free -g | head -2; df -h /dev/shm | tail -1) > gpurun_out/${TAG}_host.txt 2>&1
echo "== o64v512 x$N"; timeout 900 bash -c "$(declare -f run); N=$N; run 29511 --workload o64v512 --steps 8 --warmup 3" > gpurun_out/${TAG}_bench_o64v512_n$N.json 2> gpurun_out/${TAG}_bench_o64v512_n$N.err; tail -c 2500 gpurun_out/${TAG}_bench_o64v512_n$N.json; tail -3 gpurun_out/${TAG}_bench_o64v512_n$N.err
echo "== o100v800 x$N"; timeout 600 bash -c "$(declare -f run); N=$N; run 29512 --workload o100v800 --steps 1 --warmup 3" > gpurun_out/${TAG}_bench_o100v800_n$N.json 2> gpurun_out/${TAG}_bench_o100v800_n$N.err; tail -c 2500 gpurun_out/${TAG}_bench_o100v800_n$N.json; tail -3 gpurun_out/${TAG}_bench_o100v800_n$N.err
echo "== o40v300 x$N"; timeout 300 bash -c "$(declare -f run); N=$N; run 29513 --steps 8 --warmup 3" > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; tail -c 2500 gpurun_out/${TAG}_bench_n$N.json; tail -3 gpurun_out/${TAG}_bench_n$N.err
cat gpurun_out/${TAG}_host.txt
