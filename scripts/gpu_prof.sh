#!/bin/bash
# tests + timing + one ncu --set full capture of the fused kernel (148 triples at o=40,v=300)
TAG=${1:-r01f}
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== timing"; timeout 600 python scripts/quick_timing.py order=1 2>&1 | tail -4
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/${TAG}_fused -f python scripts/prof_run.py 5000 148 1 2>&1 | tail -3
python tools/ncu_summary.py gpurun_out/${TAG}_fused.ncu-rep > gpurun_out/${TAG}_fused_summary.txt 2>&1; head -16 gpurun_out/${TAG}_fused_summary.txt
