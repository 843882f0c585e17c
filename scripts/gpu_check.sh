#!/bin/bash
# GPU box: engine self test + the GPU parity suite; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( cd voicecontrolledrobot-var_b200/csrc
  for args in "kmajor" "mn 1"; do
    tag=$(echo $args | tr ' ' '_')
    timeout -s KILL 120 ./selftest $args > ../../gpurun_out/selftest_$tag.log 2>&1
    echo "== selftest $args -> exit $?"
    grep -E "FAIL|rc=" ../../gpurun_out/selftest_$tag.log | head -10
  done )
timeout -s KILL 900 python -m pytest tests -m gpu -q -s --timeout 600 "$@" > gpurun_out/pytest_gpu.log 2>&1
echo "== pytest exit $?"
tail -n 60 gpurun_out/pytest_gpu.log
