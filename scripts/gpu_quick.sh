#!/bin/bash
# Quick GPU pass: tensor-core tests + fp16 bench.  Usage: scripts/gpu_quick.sh tag [bench args]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > "$OUT/t_tc.log" 2>&1; echo "tc exit $?"; grep -E "^\[parity|^\[stage|passed|failed|^FAILED|conv_tc" "$OUT/t_tc.log" | head -40
timeout 900 python bench.py --steps 50 --warmup 3 --profile-out "$OUT/prof_fp16.json" "$@" > "$OUT/bench_fp16.log" 2>&1; echo "bench exit $?"; tail -2 "$OUT/bench_fp16.log" | cut -c1-600
python - <<PY
import json
d=json.load(open('$OUT/prof_fp16.json'))
print('step ms', d['step_ms_profiled'], 'kp/img', d['keypoints_per_image'])
for r in d['per_kernel']:
    print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
