#!/bin/bash
# e2e (host-buffer) rate for several pipeline chunk sizes.  Usage: scripts/gpu_e2e.sh tag
TAG=${1:-e2e}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for CH in 4 8 16 32; do
  SPB200_HOST_CHUNK=$CH timeout -s KILL 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > "$OUT/bench_$CH.log" 2>&1
  python - <<PY
import json
for l in open('$OUT/bench_$CH.log'):
    if l.startswith('{'):
        d=json.loads(l); print('chunk $CH: value %.0f  e2e %.0f img/s  (%.2f ms per 64)' % (d['value'], d['e2e']['value'], 64e3/d['e2e']['value']))
PY
done
