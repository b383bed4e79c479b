#!/bin/bash
# GPU-box check of the fused loss path: tests, then A/B bench (fused vs separate loss kernels)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py -x -q -p no:cacheprovider > gpurun_out/loss_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/loss_tests.log
for tag in fused separate; do
  if [ $tag = separate ]; then export DAD_BENCH_SEPARATE_LOSSES=1; fi
  timeout 300 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/bench_loss_$tag.json 2> gpurun_out/bench_loss_$tag.err; echo "bench $tag rc=$?"
  python - <<PYEOF
import json
d = json.load(open("gpurun_out/bench_loss_$tag.json"))
print("$tag", round(d["value"], 1), "img/s", round(d["ms_per_step"], 3), "ms; loss", d["kernel_breakdown"]["loss"], "elementwise", round(d["kernel_breakdown"]["elementwise"]["ms_per_step"], 3), "losses", d["config"]["losses"], d["clocks"]["sm_mhz"])
PYEOF
done
