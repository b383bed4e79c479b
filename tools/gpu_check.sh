#!/bin/bash
# GPU-box check (run under gpurun from the repo root): full -m gpu suite, then the default bench line.
TAG=${1:-chk}
mkdir -p gpurun_out
(time timeout 800 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_$TAG.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<EOF
import json
d = json.load(open("gpurun_out/bench_$TAG.json"))
print(d["value"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["roofline"]["frac"])
EOF
