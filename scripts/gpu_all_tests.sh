#!/bin/bash
TAG=${1:-all}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 2400 python -m pytest tests -m gpu -q -x > "$OUT/t_gpu.log" 2>&1; echo "gpu tests exit $?"; tail -5 "$OUT/t_gpu.log"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$OUT/smoke.log"
