#!/bin/bash
# Multi-GPU evidence (one box, 8 GPUs): weak scaling N = 1, 2, 4, 8; strong scaling of 512 images (BASELINE configs[3]);
# 1088x1920 top-k 2048 batch 16 per GPU at 8 GPUs (configs[4]).  Usage: scripts/gpu_multi_r02.sh tag
TAG=${1:-multi}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
run() {  # n, name, args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"; fi
  echo "$name exit $?"; tail -1 "$OUT/$name.log" | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('   value %.0f img/s  ms/step %.4f  e2e %.0f  n_gpus %d  scaling %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['n_gpus'], d['scaling']))
except Exception as ex:
    print('   no line', ex)
"
}
for N in 1 2 4 8; do run $N weak_$N --steps 100 --warmup 5 --no-cpu-baseline --no-extras; done
for N in 1 2 4 8; do run $N strong512_$N --steps 100 --warmup 5 --no-cpu-baseline --no-extras --total-batch 512; done
run 8 c4_1080p_8 --steps 100 --warmup 5 --no-cpu-baseline --no-extras --height 1088 --width 1920 --top-k 2048 --batch 16
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
