#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root, ~5 min):
#   1. the experimental attention kernel (DAD_ATT_VARIANT=4): kernel tests + device time against variants 2 / 3
#   2. headline bench with the default and, if (1) passed, with variant 4
# Everything lands in gpurun_out/round2_first/.
set -u
OUT=gpurun_out/round2_first
mkdir -p $OUT
echo "== attention kernel tests incl. experimental variant 4" | tee $OUT/summary.txt
DAD_TEST_EXPERIMENTAL=1 timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -k attention -q --timeout 90 -p no:cacheprovider \
    > $OUT/attn_tests.log 2>&1
echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -5 $OUT/attn_tests.log | tee -a $OUT/summary.txt
echo "== device time, ViT-L 518^2 B=32 (scale 0.5 = the benchmark's regime; 1.5 = rescale-heavy)" | tee -a $OUT/summary.txt
for v in 2 3 4; do
    for s in 0.5 1.5; do
        DAD_ATT_VARIANT=$v timeout -s KILL 40 python tests/gpu_attn_time.py 32 1370 16 $s 2>&1 | tail -1 | tee -a $OUT/summary.txt
    done
done
echo "== headline bench, default attention" | tee -a $OUT/summary.txt
timeout -s KILL 120 python bench.py --no-cpu-baseline > $OUT/bench_default.json 2> $OUT/bench_default.err
echo "rc=$?" | tee -a $OUT/summary.txt
echo "== headline bench, DAD_ATT_VARIANT=4 (meaningful only if the tests above passed)" | tee -a $OUT/summary.txt
DAD_ATT_VARIANT=4 timeout -s KILL 120 python bench.py --no-cpu-baseline > $OUT/bench_variant4.json 2> $OUT/bench_variant4.err
echo "rc=$?" | tee -a $OUT/summary.txt
python - <<'PY' | tee -a $OUT/summary.txt
import json
for tag in ("default", "variant4"):
    try:
        d = json.loads(open(f"gpurun_out/round2_first/bench_{tag}.json").read().strip().splitlines()[-1])
        kb = d["kernel_breakdown"]
        print(tag, round(d["value"], 1), "img/s", round(d["ms_per_step"], 2), "ms; attention", round(kb["attention"]["ms_per_step"], 2),
              "ms; gemm", round(kb["gemm_tc"]["ms_per_step"], 2), "ms; clocks", d["clocks"]["sm_mhz"])
    except Exception as ex:
        print(tag, "no record:", ex)
PY
