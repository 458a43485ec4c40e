#!/bin/bash
# Round-2 evidence pass (one GPU): all GPU tests, the full bench line with profile, the reference arm, BASELINE configs 1 and
# the 1080p frame size on one GPU, B = 1 latency, ncu launch list and full-set capture of one step.  Usage: scripts/gpu_evidence_r02.sh tag
TAG=${1:-ev2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
RX="stem_planes|planes_kernel|block_tc|halo_tc|heatmap|nms_|sample_desc|stem_tc"
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,driver_version --format=csv > "$OUT/gpu.txt" 2>&1
timeout 2400 python -m pytest tests -m gpu -q > "$OUT/t_gpu.log" 2>&1; echo "gpu tests exit $?"; tail -2 "$OUT/t_gpu.log"
timeout 900 python bench.py --steps 300 --warmup 5 --profile-out "$OUT/prof_fp16.json" > "$OUT/bench.log" 2>"$OUT/bench.err"; echo "bench exit $?"; tail -1 "$OUT/bench.log" | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.log" 2>&1; echo "ref exit $?"; tail -1 "$OUT/bench_ref.log" | cut -c1-200
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras --detector-only --batch 32 --height 240 --width 320 > "$OUT/bench_config1.log" 2>&1; echo "config1 exit $?"; tail -1 "$OUT/bench_config1.log" | cut -c1-200
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --batch 16 --height 1088 --width 1920 --top-k 2048 > "$OUT/bench_1080p_1gpu.log" 2>&1; echo "1080p exit $?"; tail -1 "$OUT/bench_1080p_1gpu.log" | cut -c1-200
timeout 600 python bench.py --steps 300 --warmup 5 --no-cpu-baseline --no-extras --batch 1 --height 240 --width 320 > "$OUT/bench_b1_240x320.log" 2>&1; echo "b1 exit $?"; tail -1 "$OUT/bench_b1_240x320.log" | cut -c1-200
for P in fp16+layer1 fp16+all; do
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-extras --precision $P --profile-out "$OUT/prof_$P.json" > "$OUT/bench_$P.log" 2>&1; echo "bench $P exit $?"; tail -1 "$OUT/bench_$P.log" | cut -c1-160
done
python scripts/ncu_target.py 3 > "$OUT/plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RX" -s 20 -c 20 --csv --log-file "$OUT/launches.csv" python scripts/ncu_target.py 3 > "$OUT/ncu1.log" 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 20 -c 20 -o "$OUT/prof" -f python scripts/ncu_target.py 3 > "$OUT/ncu2.log" 2>&1
echo "full set exit $?"
python scripts/ncu_summary.py "$OUT/prof.ncu-rep" > "$OUT/ncu_full_summary.txt" 2>&1; tail -25 "$OUT/ncu_full_summary.txt"
ls -la "$OUT"
