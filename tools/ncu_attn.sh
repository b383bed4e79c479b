#!/bin/bash
# ncu --set full + source page of ONE attention launch (run under gpurun).  usage: bash tools/ncu_attn.sh <tag> [env...]
set -u
TAG=${1:-a}
OUT=gpurun_out/ncu_attn_$TAG
mkdir -p $OUT
CMD="python tests/gpu_attn_time.py 32 1370 16 0.5"
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
tail -1 $OUT/plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 5 -c 1 -f -o $OUT/prof $CMD > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/ncu.log
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/prof.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ls -la $OUT
