#!/bin/bash
# GPU tests + bench with profile (no e2e sweep).  Usage: scripts/gpu_mid.sh tag
TAG=${1:-mid}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 1200 python -m pytest tests -m gpu -q -x > "$OUT/t_gpu.log" 2>&1; echo "gpu tests exit $?"; tail -3 "$OUT/t_gpu.log"
timeout 900 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_fp16.json" > "$OUT/bench_fp16.log" 2>&1; echo "bench exit $?"; tail -1 "$OUT/bench_fp16.log" | cut -c1-200
python - <<PY
import json
d=json.load(open('$OUT/prof_fp16.json'))
print('step ms', d['step_ms_profiled'], 'kp/img', d['keypoints_per_image'])
for r in d['per_kernel']:
    print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
