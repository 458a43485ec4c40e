#!/bin/bash
# GPU pass for the haloed-tile kernel: unit convs (both descriptor base-offset hypotheses), then the tensor-core
# suite and a bench.  Usage: scripts/gpu_halo.sh tag
TAG=${1:-halo}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
for BO in 0 1; do
  SPB200_HALO_BASEOFF=$BO timeout -s KILL 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "persistent" > "$OUT/t_halo_bo$BO.log" 2>&1
  echo "halo unit (base_off=$BO) exit $?"; grep -E "passed|failed|^\[conv kernel" "$OUT/t_halo_bo$BO.log" | cut -c1-220 | head -12
done
for OR in 0 1; do
  SPB200_HALO_ORIENT=$OR timeout -s KILL 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -s -k "persistent and fp16" > "$OUT/t_halo_or$OR.log" 2>&1
  echo "halo unit (orient=$OR) exit $?"; grep -E "passed|failed|^\[conv kernel" "$OUT/t_halo_or$OR.log" | cut -c1-220 | head -8
done
timeout -s KILL 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s > "$OUT/t_tc.log" 2>&1; echo "tc exit $?"; grep -E "^\[parity|^\[stage|passed|failed|^FAILED|conv kernel" "$OUT/t_tc.log" | cut -c1-200 | head -50
timeout -s KILL 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_fp16.json" > "$OUT/bench_fp16.log" 2>&1; echo "bench exit $?"; tail -2 "$OUT/bench_fp16.log" | cut -c1-400
SPB200_NO_HALO=1 timeout -s KILL 600 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --profile-out "$OUT/prof_fp16_nohalo.json" > "$OUT/bench_fp16_nohalo.log" 2>&1; echo "bench(no halo) exit $?"
python - <<PY
import json
for f in ('prof_fp16.json', 'prof_fp16_nohalo.json'):
    try:
        d=json.load(open('$OUT/'+f))
    except Exception as ex:
        print(f, 'missing', ex); continue
    print(f, 'step ms', d['step_ms_profiled'], 'kp/img', d['keypoints_per_image'])
    for r in d['per_kernel']:
        print('%-36s %7.3f ms %5.1f%%  %s %s' % (r['kernel'], r['ms'], 100*r['share'], ('%.0f TF/s (%.1f%%)' % (r['tflops'], 100*r['frac_tc_sustained'])) if 'tflops' in r else '', ('%.0f GB/s (%.1f%%)' % (r['gbs'], 100*r['frac_hbm'])) if 'gbs' in r else ''))
PY
