#!/bin/bash
# iteration call: parity + quick timing (+ optional ncu capture of a 148-triple range)
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
if [ -n "$SANITIZE" ]; then
echo "== memcheck smoke"; timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
fi
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== timing"; timeout 600 python scripts/quick_timing.py $VARIANTS 2>&1 | tail -16
if [ -n "$NCU" ]; then
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:pt_fused -c 1 -o gpurun_out/prof_fused_$NCU -f python scripts/prof_run.py 5000 148 1 2>&1 | tail -3
fi
