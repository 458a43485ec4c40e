#!/bin/bash
# Fused detector tail: identity check against the logits path, GPU tests, bench profile.  Usage: scripts/gpu_heat.sh tag [all]
TAG=${1:-heat}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python scripts/fused_heat_check.py > "$OUT/check.log" 2>&1; echo "check exit $?"; tail -6 "$OUT/check.log"
if [ "$2" = "all" ]; then
  timeout 2400 python -m pytest tests -m gpu -q -x > "$OUT/t_all.log" 2>&1; echo "gpu tests exit $?"; tail -4 "$OUT/t_all.log"
else
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -m gpu -q -x > "$OUT/t.log" 2>&1; echo "tests exit $?"; tail -4 "$OUT/t.log"
fi
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --profile-out "$OUT/prof_fp16.json" > "$OUT/bench_fp16.log" 2>&1; echo "bench exit $?"; tail -1 "$OUT/bench_fp16.log" | cut -c1-200
python - <<PY
import json
d=json.load(open('$OUT/prof_fp16.json'))
print('step ms', d['step_ms_profiled'], 'kp/img', d['keypoints_per_image'])
for r in d['per_kernel']:
    print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
