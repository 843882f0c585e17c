#!/bin/bash
# A/B of the 16-bit GRU input projection: GPU parity suite, then the headline bench with and without it.
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "== pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
for v in 1 0; do export VAR_EPI_TMA=$v
  timeout -s KILL 600 python bench.py --workload ithor_b256 --steps 20 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-reward > gpurun_out/bench_x16_$v.json 2> gpurun_out/bench_x16_$v.err
  echo "== x16=$v exit $?"; python - <<E
import json
try:
    d=json.loads(open("gpurun_out/bench_x16_$v.json").read().strip().splitlines()[-1])
    print("x16=$v", d["ms_per_step"], d["e2e"]["ms_per_step"])
    for k,x in d["kernels"].items(): print("   ",k,x["ms_per_step"],x.get("tflops"))
except Exception as e: print("parse failed", e); print(open("gpurun_out/bench_x16_$v.err").read()[-1500:])
E
done
