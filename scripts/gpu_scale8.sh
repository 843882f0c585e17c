#!/bin/bash
# 8-GPU evidence: DP test, iTHOR weak-scaling line (with sharded reward queries), Kuka global-batch-8192 lines at N=8 and N=4.
mkdir -p gpurun_out
# (the 2-GPU data-parallel tests run on the 2-GPU box: scripts/gpu_multi.sh)
run() {  # N workload extra-args
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus $1 --steps 20 --warmup 5 --workload $2 $3 > gpurun_out/bench_$2_n$1.json 2> gpurun_out/bench_$2_n$1.err
  echo "== bench $2 N=$1 exit $?"; tail -n 2 gpurun_out/bench_$2_n$1.err | cut -c1-200; head -c 260 gpurun_out/bench_$2_n$1.json; echo
}
run 8 ithor_b256 "--no-cpu-baseline --no-torch-baseline"
run 8 kuka_dp8192 "--no-cpu-baseline --no-torch-baseline --no-reward"
run 4 kuka_dp8192 "--no-cpu-baseline --no-torch-baseline --no-reward"
