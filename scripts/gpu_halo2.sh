#!/bin/bash
TAG=${1:-halo}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout -s KILL 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "persistent" > "$OUT/t_halo.log" 2>&1
echo "halo unit exit $?"; grep -E "passed|failed|^\[conv kernel" "$OUT/t_halo.log" | cut -c1-220 | head -12
SPB200_HALO_STATS=1 timeout -s KILL 300 python scripts/ncu_target.py 1 2>&1 | cut -c1-420 | tail -16
timeout -s KILL 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > "$OUT/t_tc.log" 2>&1; echo "tc exit $?"; grep -E "^\[parity|passed|failed|^FAILED|conv kernel" "$OUT/t_tc.log" | cut -c1-200 | head -30
bash scripts/gpu_ab.sh $TAG "A=1"
