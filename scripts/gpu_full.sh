#!/bin/bash
# Tensor-core + parity suites and a profiled bench.  Usage: scripts/gpu_full.sh tag
TAG=${1:-full}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout -s KILL 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > "$OUT/t_tc.log" 2>&1; echo "tc exit $?"; grep -E "^\[parity|^\[stage|passed|failed|^FAILED|conv kernel" "$OUT/t_tc.log" | cut -c1-200 | head -40
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > "$OUT/t_parity.log" 2>&1; echo "parity exit $?"; tail -3 "$OUT/t_parity.log"
bash scripts/gpu_ab.sh $TAG "A=1"
