#!/bin/bash
# GPU-box check (run under gpurun from the repo root): full -m gpu suite, smoke, the default bench line, attention A/B.
TAG=${1:-chk}
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8) > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PYEOF
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], d["clocks"], "roofline", d["roofline"]["frac"])
for k, v in d["kernel_breakdown"].items():
    print(" ", k, round(v["ms_per_step"], 3), "ms", v["launches_per_step"], "launches", round(v["frac"], 3))
print("eager", d.get("gpu_eager_baseline"))
print("cpu", d.get("cpu_baseline"))
PYEOF
for w in train train4; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline --no-gpu-eager --steps 8 --warmup 3 > gpurun_out/bench_${w}_$TAG.json 2> gpurun_out/bench_${w}_$TAG.err
  python -c "import json; d = json.load(open('gpurun_out/bench_${w}_$TAG.json')); print('$w', round(d['value'], 1), d['unit'], round(d['ms_per_step'], 2), 'ms/step', {k: round(v['ms_per_step'], 2) for k, v in d.get('kernel_breakdown', {}).items()})"
done
SCALES=0.5 POLYS="2" bash tools/attn_ab.sh > /dev/null 2>&1; tail -4 gpurun_out/attn_ab/summary.txt
