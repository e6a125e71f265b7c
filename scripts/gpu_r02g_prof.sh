#!/bin/bash
# round 2 profiling evidence (one GPU): ncu launch list of a short bench run, --set full captures of the
# fused (T) kernel and of the integrals-from-vertex GEMM, DRAM traffic of one bench step.
TAG=${1:-r02g}
mkdir -p gpurun_out
nvidia-smi > gpurun_out/${TAG}_nvidia_smi.txt 2>&1
echo "== launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/${TAG}_bench_under_ncu.log; wc -l gpurun_out/${TAG}_launches.csv
echo "== ncu full fused"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/${TAG}_fused -f python scripts/prof_run.py 5000 148 1 2>&1 | tail -2
python tools/ncu_summary.py gpurun_out/${TAG}_fused.ncu-rep > gpurun_out/${TAG}_fused_summary.txt 2>&1; head -16 gpurun_out/${TAG}_fused_summary.txt
echo "== dram traffic of one bench step"; timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:pt_fused -c 1 --csv --log-file gpurun_out/${TAG}_step_traffic.csv python scripts/prof_run.py step 3 1 2>&1 | tail -1; tail -4 gpurun_out/${TAG}_step_traffic.csv | cut -d, -f13-
echo "== ncu full vertex gemm"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:vertex_gemm -s 1 -c 1 -o gpurun_out/${TAG}_vertex_gemm -f python scripts/prof_vertex.py 24 512 1024 2>&1 | tail -2
python tools/ncu_summary.py gpurun_out/${TAG}_vertex_gemm.ncu-rep > gpurun_out/${TAG}_vertex_gemm_summary.txt 2>&1; head -16 gpurun_out/${TAG}_vertex_gemm_summary.txt
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json
rm -f gpurun_out/${TAG}_fused.ncu-rep.tmp
