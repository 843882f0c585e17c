#!/bin/bash
# GPU box: smoke + the benchmark lines.  Usage: bash scripts/gpu_bench.sh [extra bench args]
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 3 gpurun_out/smoke.log
for wl in kuka_b64 ithor_b256; do
  timeout -s KILL 900 python bench.py --workload $wl "$@" > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
  echo "== bench $wl exit $?"; tail -n 5 gpurun_out/bench_$wl.err; head -c 3000 gpurun_out/bench_$wl.json; echo
done
