#!/bin/bash
# Full-set ncu capture of the kernels matching a regex in one step of the bench workload.
# Usage: scripts/gpu_ncu_one.sh tag regex [skip] [count]
TAG=${1:-ncu1}; RX=${2:-stem_planes}; SKIP=${3:-1}; CNT=${4:-1}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python scripts/ncu_target.py 3 > "$OUT/plain.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT -o "$OUT/prof" -f python scripts/ncu_target.py 3 > "$OUT/ncu.log" 2>&1
echo "ncu exit $?"; ls -la "$OUT"
