#!/bin/bash
# Runs the engine self test on the GPU box; each mode in its own process so a
# faulting hypothesis cannot poison the others.
mkdir -p gpurun_out
cd "voicecontrolledrobot-var_b200/csrc"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > ../../gpurun_out/gpu.txt 2>&1
for args in "kmajor" "mn 1" "mn 2" "mn 3" "mn 4" "mn 5" "perf"; do
  tag=$(echo $args | tr ' ' '_')
  timeout -s KILL 120 ./selftest $args > ../../gpurun_out/selftest_$tag.log 2>&1
  echo "== selftest $args -> exit $?" | tee -a ../../gpurun_out/selftest_summary.log
  tail -n 40 ../../gpurun_out/selftest_$tag.log
done
