#!/bin/bash
# round 2, call b: the new bench at the three headline shapes on one GPU
mkdir -p gpurun_out
TAG=${1:-r02b}
echo "== bench o40v300"; timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== bench o64v512"; timeout 900 python bench.py --workload o64v512 --steps 4 --e2e-share 24 > gpurun_out/${TAG}_bench_o64v512.json 2> gpurun_out/${TAG}_bench_o64v512.err; tail -c 3000 gpurun_out/${TAG}_bench_o64v512.json; tail -5 gpurun_out/${TAG}_bench_o64v512.err
echo "== bench o100v800"; timeout 900 python bench.py --workload o100v800 > gpurun_out/${TAG}_bench_o100v800.json 2> gpurun_out/${TAG}_bench_o100v800.err; tail -c 3000 gpurun_out/${TAG}_bench_o100v800.json; tail -5 gpurun_out/${TAG}_bench_o100v800.err
free -g | head -2; df -h /dev/shm | tail -1; nproc
