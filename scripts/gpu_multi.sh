#!/bin/bash
# N-GPU box: NCCL data-parallel test + bench lines under torchrun.  Usage: bash scripts/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "== dist pytest exit $?"; tail -n 4 gpurun_out/pytest_dist.log
for wl in ithor_b256 kuka_dp8192; do
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu-baseline --no-reward > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  echo "== bench $wl N=$N exit $?"; tail -n 3 gpurun_out/bench_${wl}_n$N.err; head -c 600 gpurun_out/bench_${wl}_n$N.json; echo
done
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload kuka_dp8192 --no-cpu-baseline --no-reward > gpurun_out/bench_kuka_dp8192_n1.json 2> gpurun_out/bench_kuka_dp8192_n1.err
echo "== bench kuka_dp8192 N=1 exit $?"; head -c 600 gpurun_out/bench_kuka_dp8192_n1.json; echo
