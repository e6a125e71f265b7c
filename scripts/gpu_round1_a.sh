#!/bin/bash
# Round-1 evidence call: correctness tiers, FP64 ceilings, bench line, ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
(nproc; free -g; lscpu | head -25) > gpurun_out/host.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
echo "== microbench"; timeout 300 python tools/fp64_microbench.py > gpurun_out/fp64_microbench.json 2>gpurun_out/microbench.err; tail -c 2500 gpurun_out/fp64_microbench.json
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_r01a.json 2> gpurun_out/bench_r01a.err; tail -c 4000 gpurun_out/bench_r01a.json; tail -5 gpurun_out/bench_r01a.err
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/prof_fused_r01a -f python scripts/prof_run.py 5000 16 2>&1 | tail -4
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01a.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/launches_r01a.csv
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
