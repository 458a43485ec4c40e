#!/bin/bash
TAG=${1:-split}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout -s KILL 300 python scripts/split_check.py 240 320 > "$OUT/split_240.log" 2>&1; echo "split 240 exit $?"; tail -8 "$OUT/split_240.log"
timeout -s KILL 300 python scripts/split_check.py 208 272 > "$OUT/split_208.log" 2>&1; echo "split 208 exit $?"; tail -5 "$OUT/split_208.log"
timeout -s KILL 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > "$OUT/t_tc.log" 2>&1; echo "tc tests exit $?"; tail -3 "$OUT/t_tc.log"
