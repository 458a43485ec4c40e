#!/bin/bash
# Suites + the default bench line (with the CPU baseline) + the reference arm.  Usage: scripts/gpu_bench.sh tag
TAG=${1:-bench}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout -s KILL 900 python -m pytest tests -m gpu -q -x > "$OUT/t_gpu.log" 2>&1; echo "gpu tests exit $?"; tail -3 "$OUT/t_gpu.log"
timeout -s KILL 600 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$OUT/smoke.log"
timeout -s KILL 900 python bench.py --profile-out "$OUT/prof.json" > "$OUT/bench.log" 2>&1; echo "bench exit $?"; tail -1 "$OUT/bench.log"
timeout -s KILL 600 python bench.py --impl reference --steps 5 --warmup 1 > "$OUT/bench_ref.log" 2>&1; echo "ref exit $?"; tail -1 "$OUT/bench_ref.log" | cut -c1-300
