#!/bin/bash
# First run of the CTA-pair kernel (SPB200_PAIR=1): equality with the default path, the tensor-core tests, per-kernel times.
TAG=${1:-pair}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > "$OUT/gpu.txt" 2>&1
timeout -s KILL 120 python scripts/pair_check.py 2 240 320 > "$OUT/pair_small.log" 2>&1; echo "pair small exit $?"; tail -5 "$OUT/pair_small.log"
timeout -s KILL 180 python scripts/pair_check.py 64 480 640 > "$OUT/pair_big.log" 2>&1; echo "pair big exit $?"; tail -5 "$OUT/pair_big.log"
timeout 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_default.json" > "$OUT/bench_default.log" 2>&1; echo "bench default exit $?"; tail -1 "$OUT/bench_default.log" | cut -c1-200
SPB200_PAIR=1 timeout -s KILL 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_pair.json" > "$OUT/bench_pair.log" 2>&1; echo "bench pair exit $?"; tail -1 "$OUT/bench_pair.log" | cut -c1-200
SPB200_PAIR=1 timeout -s KILL 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x > "$OUT/t_pair.log" 2>&1; echo "pair tests exit $?"; tail -3 "$OUT/t_pair.log"
python - <<PY
import json
for n in ('default','pair'):
    try:
        d=json.load(open('$OUT/prof_%s.json' % n))
    except Exception as ex:
        print(n, 'no profile', ex); continue
    print(n, 'step ms', d['step_ms_profiled'])
    for r in d['per_kernel']:
        print('  %-36s %7.3f ms %s' % (r['kernel'], r['ms'], ('%.1f%%' % (100*r['frac_tc_sustained'])) if 'tflops' in r else ''))
PY
