#!/bin/bash
# Training-path check (run under gpurun): kernel + backward parity tests, then the student train step with the three weight-gradient
# operand paths (DAD_WGRAD_PATH: 2 MN-major, 1 padded-space copies, 0 im2col^T + transposes) on the same box.
OUT=gpurun_out/train_ab
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_model_backward.py tests/test_gpu_kernels.py -k "not attention" -q -p no:cacheprovider 2>&1 | tail -5 | tee $OUT/tests.log
for v in ${PATHS:-2 1 0 2}; do
  export DAD_WGRAD_PATH=$v
  timeout 300 python bench.py --workload train --no-cpu-baseline --no-gpu-eager --steps 8 --warmup 3 > $OUT/bench_$v.json 2> $OUT/bench_$v.err || tail -3 $OUT/bench_$v.err
  python - <<PYEOF
import json
d = json.load(open("$OUT/bench_$v.json"))
print("wgrad=$v: %.1f img/s %.2f ms/step" % (d["value"], d["ms_per_step"]), {k: round(x["ms_per_step"], 2) for k, x in d.get("kernel_breakdown", {}).items()})
PYEOF
done | tee $OUT/summary.txt
