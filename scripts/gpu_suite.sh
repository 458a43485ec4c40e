#!/bin/bash
# Full GPU pass on a B200 box (run through gpurun): parity tests, smoke, bench.  Logs -> gpurun_out/.
# Usage: scripts/gpu_suite.sh [tag]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi > "$OUT/nvidia-smi.txt" 2>&1
nproc > "$OUT/nproc.txt"
python -c "import __graft_entry__ as g; g.build()" > "$OUT/build.log" 2>&1 || echo "BUILD FAILED"
echo "== parity (fp32 + stage kernels)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x > "$OUT/t_parity.log" 2>&1; echo "exit $?"; tail -5 "$OUT/t_parity.log"
echo "== tensor-core tests"
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > "$OUT/t_tc.log" 2>&1; echo "exit $?"; tail -5 "$OUT/t_tc.log"
echo "== smoke"
timeout 300 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "exit $?"; tail -3 "$OUT/smoke.log"
echo "== bench fp16"
timeout 900 python bench.py --steps 10 --warmup 3 --profile-out "$OUT/prof_fp16.json" > "$OUT/bench_fp16.log" 2>&1; echo "exit $?"; tail -2 "$OUT/bench_fp16.log"
echo "== bench fp32"
timeout 900 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_fp32.json" > "$OUT/bench_fp32.log" 2>&1; echo "exit $?"; tail -2 "$OUT/bench_fp32.log"
