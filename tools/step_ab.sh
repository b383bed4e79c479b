#!/bin/bash
# Same-box A/B of the whole bench step with different attention kernels (run under gpurun): is the step time- or energy-bound?
OUT=gpurun_out/step_ab
mkdir -p $OUT
for round in 1 2; do
  for v in ${VARIANTS:-5 2}; do
    DAD_ATT_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager --steps 24 --warmup 4 > $OUT/bench_v${v}_$round.json 2> $OUT/bench_v${v}_$round.err
    python - <<PYEOF
import json
d = json.load(open("$OUT/bench_v${v}_$round.json"))
kb = d["kernel_breakdown"]
print("variant $v run $round: %.1f img/s %.2f ms/step e2e %.1f sm %s MHz attention %.2f ms gemm %.2f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], kb["attention"]["ms_per_step"], kb["gemm_tc"]["ms_per_step"]))
PYEOF
  done
done | tee $OUT/summary.txt
