#!/bin/bash
# A/B of the bucketed all-reduce on N GPUs (device-timed `value` only).
N=${1:-2}
mkdir -p gpurun_out
for b in 1 0; do
  VAR_DP_BUCKET=$b timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-reward --no-cpu-baseline --no-torch-baseline > gpurun_out/dp_bucket${b}_n$N.json 2> gpurun_out/dp_bucket${b}_n$N.err
  echo "== bucket=$b N=$N exit $?"; python - <<P
import json
d=json.loads(open("gpurun_out/dp_bucket${b}_n$N.json").read().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3))
P
done
