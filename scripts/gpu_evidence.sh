#!/bin/bash
# Evidence call: correctness tiers, bench line, ncu launch list + full capture + DRAM traffic of one bench step.
# usage: scripts/gpu_evidence.sh <tag>   (outputs gpurun_out/<tag>_*)
TAG=${1:-r01d}
mkdir -p gpurun_out
nvidia-smi > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12
echo "== bench"; timeout 1200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 4000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/${TAG}_fused -f python scripts/prof_run.py 5000 148 1 2>&1 | tail -3
python tools/ncu_summary.py gpurun_out/${TAG}_fused.ncu-rep > gpurun_out/${TAG}_fused_summary.txt 2>&1; head -20 gpurun_out/${TAG}_fused_summary.txt
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; tail -2 gpurun_out/${TAG}_bench_under_ncu.log; wc -l gpurun_out/${TAG}_launches.csv
echo "== dram traffic of one bench step"; timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pt_fused -c 1 --csv --log-file gpurun_out/${TAG}_step_traffic.csv python scripts/prof_run.py step 3 1 2>&1 | tail -1; tail -4 gpurun_out/${TAG}_step_traffic.csv | cut -d, -f13-
echo "== feed ceiling"; timeout 600 python scripts/feed_ceiling.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_feed.txt
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
