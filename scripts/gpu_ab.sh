#!/bin/bash
# A/B bench runs under different environment settings.  Usage: scripts/gpu_ab.sh tag "ENV1=.. ENV2=.." "ENV.." ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
i=0
for CFG in "$@"; do
  i=$((i+1))
  env $CFG timeout -s KILL 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_$i.json" > "$OUT/bench_$i.log" 2>&1; echo "[$CFG] bench exit $?"
  python - <<PY
import json
d=json.load(open('$OUT/prof_$i.json'))
print('step ms %.3f' % d['step_ms_profiled'])
for r in d['per_kernel']:
    print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
done
