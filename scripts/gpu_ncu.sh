#!/bin/bash
# ncu evidence (one GPU): launch list of one step + full-set capture of one step.  Usage: scripts/gpu_ncu.sh tag
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
python scripts/ncu_target.py 3 > "$OUT/plain.log" 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:block_tc|halo_tc|heatmap|nms_|sample_desc|sort_emit|stem_tc" -s 20 -c 20 --csv --log-file "$OUT/launches.csv" python scripts/ncu_target.py 3 > "$OUT/ncu1.log" 2>&1
echo "launch list exit $?"
python scripts/ncu_target.py 3 > "$OUT/plain2.log" 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:block_tc|halo_tc|heatmap|nms_|sample_desc|sort_emit|stem_tc" -s 20 -c 20 -o "$OUT/prof" -f python scripts/ncu_target.py 3 > "$OUT/ncu2.log" 2>&1
echo "full set exit $?"; ls -la "$OUT"
