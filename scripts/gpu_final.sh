#!/bin/bash
# Final evidence run: smoke, full bench lines (with cpu baseline + reward) for the headline and
# side workloads, reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 1 gpurun_out/smoke.log
for wl in ithor_b256 kuka_b64; do
  timeout -s KILL 600 python bench.py --workload $wl > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err; echo "== bench $wl exit $?"
done
for wl in kuka_dp8192 mfcc_4s; do
  timeout -s KILL 600 python bench.py --workload $wl --no-cpu-baseline --no-reward --steps 10 > gpurun_out/final_$wl.json 2> gpurun_out/final_$wl.err; echo "== bench $wl exit $?"
done
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "== reference exit $?"
python scripts/profile_mfcc.py > gpurun_out/final_mfcc.log 2>&1; cat gpurun_out/final_mfcc.log
for f in gpurun_out/final_*.json; do echo $f; head -c 400 $f; echo; done
