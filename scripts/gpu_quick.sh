#!/bin/bash
# Quick check call: GPU parity tests + kernel timing of (20,100) and one (40,300) bench step.
# usage: scripts/gpu_quick.sh [variant ...]   (variants = "key=value,key=value" option sets)
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
echo "== timing"; timeout 600 python scripts/quick_timing.py "$@" 2>&1 | tail -12
