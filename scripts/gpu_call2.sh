#!/bin/bash
mkdir -p gpurun_out
echo "== pytest o40"; timeout 900 python -m pytest tests -m gpu -q -k "o40" 2>&1 | tail -8
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/prof_fused_r01 -f python scripts/prof_run.py 5000 16 2>&1 | tail -4
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 3000 gpurun_out/bench_r01.json; tail -5 gpurun_out/bench_r01.err
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/launches_r01.csv
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2
