#!/bin/bash
# operator tests + launch-shape sweeps of the im2col-free conv kernels
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py -x -q 2>&1 | tail -6
timeout 200 python scripts/profile_halo.py 2>&1 | tail -24
for geo in "256 96 96 32 32 3 3 1 1 1 1" "256 48 48 32 64 3 3 1 1 1 1"; do
  for v in "VAR_HALO32=1" "VAR_HALO32=0" "VAR_HALO32_RESIDENT=0" "VAR_HALO32_CPS=3 VAR_HALO32_RESIDENT=0" "VAR_HALO32_CPS=1 VAR_HALO32_SLOTS=3"; do
    echo "== $geo | $v"; env $v timeout 100 python scripts/profile_layer.py $geo 2>&1 | grep fwd
  done
done
