#!/bin/bash
# GPU-box profiling pass (run under gpurun from the repo root):  bash tools/profile_round.sh <tag>
#   1. plain bench (must exit 0 before anything runs under ncu)
#   2. ncu launch list of the same command (device time per launch; cold-cache, serialised -> compare SHARES)
#   3. ncu --set full of one steady-state forward's kernels (DRAM traffic / tensor-pipe utilisation per launch)
# Everything lands in gpurun_out/; tools/summarize_ncu.py turns it into the tracked summaries under profiles/.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --graph 0"
$CMD > $OUT/plain_$TAG.json 2> $OUT/plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/plain_$TAG.err; exit 1; }
cat $OUT/plain_$TAG.json
# launch list: all launches (weight packing, 3+3 warm-ups, timed steps, profile pass)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$TAG.csv \
    $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$? rows=$(wc -l < $OUT/launches_$TAG.csv)"
# full-set capture: SKIP launches = packing + first forwards; then one whole forward+loss (~240 launches).
# DAD_NCU_SKIP can be tuned; the summary script keys on kernel names, not on positions.
SKIP=${DAD_NCU_SKIP:-1200}
COUNT=${DAD_NCU_COUNT:-245}
timeout 1500 ncu --set full --clock-control none -s $SKIP -c $COUNT -f -o /tmp/full_$TAG \
    $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
# the report can exceed gpurun's 64 MiB merge limit: export the raw page here, keep the .ncu-rep only if small
ncu -i /tmp/full_$TAG.ncu-rep --page raw --csv > $OUT/full_raw_$TAG.csv 2>/dev/null
sz=$(stat -c %s /tmp/full_$TAG.ncu-rep 2>/dev/null || echo 0)
[ "$sz" -gt 0 ] && [ "$sz" -lt 40000000 ] && cp /tmp/full_$TAG.ncu-rep $OUT/
echo "report bytes=$sz raw rows=$(wc -l < $OUT/full_raw_$TAG.csv)"
ls -la $OUT | tail -8
