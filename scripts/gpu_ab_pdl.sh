#!/bin/bash
TAG=${1:-pdl}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -q -x > "$OUT/t.log" 2>&1; echo "tests exit $?"; tail -3 "$OUT/t.log"
for V in 0 1; do
SPB200_NO_PDL=$V timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras --profile-out "$OUT/prof_nopdl$V.json" > "$OUT/bench_nopdl$V.log" 2>&1; echo "bench NO_PDL=$V exit $?"; python - <<PY
import json
d=json.loads(open('$OUT/bench_nopdl$V.log').read().strip().splitlines()[-1])
print('NO_PDL=$V value %.0f img/s  ms/step %.4f  e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
PY
done
SPB200_NO_PDL=0 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras --detector-only --batch 32 --height 240 --width 320 > "$OUT/bench_c1_pdl.log" 2>&1; tail -1 "$OUT/bench_c1_pdl.log" | cut -c1-140
SPB200_NO_PDL=1 timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras --detector-only --batch 32 --height 240 --width 320 > "$OUT/bench_c1_nopdl.log" 2>&1; tail -1 "$OUT/bench_c1_nopdl.log" | cut -c1-140
