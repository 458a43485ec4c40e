#!/bin/bash
TAG=${1:-sw}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -q -s > "$OUT/t_split.log" 2>&1; echo "split tests exit $?"; tail -3 "$OUT/t_split.log"
timeout 1500 python -m pytest tests/test_gpu_wide.py -m gpu -q -s > "$OUT/t_wide.log" 2>&1; echo "wide tests exit $?"; tail -4 "$OUT/t_wide.log"
