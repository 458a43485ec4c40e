#!/bin/bash
# New stem against the old one + profile.  Usage: scripts/gpu_stem.sh tag
TAG=${1:-stem}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -s -k "plane_fed" > "$OUT/t_stem.log" 2>&1; echo "stem test exit $?"; grep -E "stem planes|passed|failed|Error|error" "$OUT/t_stem.log" | head -20
timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_fp16.json" > "$OUT/bench_fp16.log" 2>&1; echo "bench exit $?"; tail -1 "$OUT/bench_fp16.log" | cut -c1-200
python - <<PY
import json
d=json.load(open('$OUT/prof_fp16.json'))
print('step ms', d['step_ms_profiled'], 'kp/img', d['keypoints_per_image'])
for r in d['per_kernel']:
    print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
