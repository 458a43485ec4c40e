#!/bin/bash
# Evidence pass for profiles/: full bench line (with CPU baseline), reference arm, event profile, ncu launch list and
# full-set capture of one step.  Usage: scripts/gpu_evidence.sh tag
TAG=${1:-ev}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
RX="stem_planes|planes_kernel|block_tc|halo_tc|heatmap|nms_|sample_desc|stem_tc"
timeout 900 python bench.py --steps 300 --warmup 5 --profile-out "$OUT/prof_fp16.json" > "$OUT/bench.log" 2>&1; echo "bench exit $?"; tail -1 "$OUT/bench.log" | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.log" 2>&1; echo "ref exit $?"; tail -1 "$OUT/bench_ref.log" | cut -c1-300
python scripts/ncu_target.py 3 > "$OUT/plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:$RX" -s 20 -c 20 --csv --log-file "$OUT/launches.csv" python scripts/ncu_target.py 3 > "$OUT/ncu1.log" 2>&1
echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k "regex:$RX" -s 20 -c 20 -o "$OUT/prof" -f python scripts/ncu_target.py 3 > "$OUT/ncu2.log" 2>&1
echo "full set exit $?"; ls -la "$OUT"
