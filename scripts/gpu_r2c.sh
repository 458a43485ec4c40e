#!/bin/bash
TAG=${1:-r2c}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -q -x -s > "$OUT/t_split.log" 2>&1; echo "split tests exit $?"; tail -5 "$OUT/t_split.log"
timeout 1500 python -m pytest tests/test_gpu_wide.py -m gpu -q -s > "$OUT/t_wide.log" 2>&1; echo "wide tests exit $?"; tail -8 "$OUT/t_wide.log"
for P in fp16+layer1 fp16+encoder fp16+all; do
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --precision $P --profile-out "$OUT/prof_$P.json" > "$OUT/bench_$P.log" 2>&1; echo "bench $P exit $?"; tail -1 "$OUT/bench_$P.log" | cut -c1-160
done
python - <<PY
import json
for n in ('fp16+layer1','fp16+encoder','fp16+all'):
    try:
        d=json.load(open('$OUT/prof_%s.json' % n))
    except Exception as ex:
        print(n, 'no profile', ex); continue
    print(n, 'step ms', d['step_ms_profiled'])
    for r in d['per_kernel']:
        print('  %-36s %7.3f ms %s' % (r['kernel'], r['ms'], ('%.1f%%' % (100*r['frac_tc_sustained'])) if 'tflops' in r else ''))
PY
