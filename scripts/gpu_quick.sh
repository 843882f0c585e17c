#!/bin/bash
# GPU parity suite + short bench lines (no CPU / torch baselines): usage scripts/gpu_quick.sh [workloads...]
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
for wl in ${@:-ithor_b256 kuka_dp8192}; do
  timeout -s KILL 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-reward > gpurun_out/quick_$wl.json 2> gpurun_out/quick_$wl.err
  python - <<E
import json
try:
    d=json.loads(open("gpurun_out/quick_$wl.json").read().strip().splitlines()[-1])
    print("$wl", round(d["ms_per_step"],3), round(d["value"]), "e2e", round(d["e2e"]["ms_per_step"],3))
    print("   ", {k:x["ms_per_step"] for k,x in d["kernels"].items() if x["ms_per_step"]>0.1})
except Exception as e: print("parse failed", e); print(open("gpurun_out/quick_$wl.err").read()[-1500:])
E
done
