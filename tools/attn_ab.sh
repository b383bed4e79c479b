#!/bin/bash
# GPU-box A/B of the attention kernels (run under gpurun from the repo root): kernel tests, then device time per variant.
set -u
OUT=gpurun_out/attn_ab
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -k attention -q -x --timeout 120 -p no:cacheprovider > $OUT/tests.log 2>&1
echo "pytest rc=$?" | tee $OUT/summary.txt
tail -5 $OUT/tests.log | tee -a $OUT/summary.txt
for cfg in "32 1370 16" "8 5477 16"; do
  for s in ${SCALES:-0.5 1.5}; do
    for p in ${POLYS:-0 2 3 4 5}; do
      for v in ${VARIANTS:-5}; do
        DAD_ATT_VARIANT=$v DAD_ATT_POLY5=$p timeout -s KILL 60 python tests/gpu_attn_time.py $cfg $s 2>&1 | tail -1 | tee -a $OUT/summary.txt
      done
    done
    DAD_ATT_VARIANT=2 timeout -s KILL 60 python tests/gpu_attn_time.py $cfg $s 2>&1 | tail -1 | tee -a $OUT/summary.txt
  done
done
