#!/bin/bash
# Multi-GPU records (run under `gpurun --gpus N`): bash tools/multi_gpu.sh N "<workloads>"
# Each bench runs under its own timeout; a hang in the clean NCCL teardown falls back to the legacy exit for the next runs.
N=${1:-2}
WL=${2:-"c5"}
OUT=gpurun_out/multi
mkdir -p $OUT
export NCCL_DEBUG=WARN
for w in $WL; do
  for attempt in 1 2; do
    timeout -s KILL 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
        bench.py --gpus $N --workload $w --steps 8 --warmup 3 > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
    rc=$?
    echo "workload $w N=$N attempt $attempt clean_exit=${DAD_BENCH_CLEAN_EXIT:-1} rc=$rc"
    if [ $rc -eq 0 ] && [ -s $OUT/bench_${w}_n$N.json ]; then break; fi
    tail -5 $OUT/bench_${w}_n$N.err
    export DAD_BENCH_CLEAN_EXIT=0
  done
  python - <<PYEOF
import json
try:
    d = json.loads(open("$OUT/bench_${w}_n$N.json").read().strip().splitlines()[-1])
    print("$w N=$N:", round(d["value"], 1), d["unit"], round(d["ms_per_step"], 2), "ms/step", "e2e", round(d["e2e"]["value"], 1), d["clocks"])
except Exception as ex:
    print("$w N=$N: no record:", ex)
PYEOF
done
if [ "$N" -ge 2 ] && [ "${RUN_PYTEST:-0}" = "1" ]; then
  timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -k "tensors_device" -p no:cacheprovider 2>&1 | tail -3
fi
