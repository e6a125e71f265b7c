#!/bin/bash
# First GPU call: correctness tiers + FP64 ceilings + first timings.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | head -20 >> gpurun_out/host.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest small"; timeout 900 python -m pytest tests -m gpu -q -k "not o40 and not o20" 2>&1 | tail -40
echo "== microbench"; timeout 300 python tools/fp64_microbench.py > gpurun_out/fp64_microbench.json 2>gpurun_out/microbench.err; tail -c 1500 gpurun_out/fp64_microbench.json
echo "== pytest o20"; timeout 600 python -m pytest tests -m gpu -q -k "o20" 2>&1 | tail -15
echo "== timing"; timeout 900 python scripts/quick_timing.py 2>&1 | tail -20
