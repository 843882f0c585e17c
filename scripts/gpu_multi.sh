#!/bin/bash
# N-GPU box: NCCL data-parallel test + bench lines under torchrun.  Usage: bash scripts/gpu_multi.sh N [workloads...]
N=${1:-2}; shift
WLS=${@:-ithor_b256 kuka_dp8192}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/pytest_dist.log 2>&1; echo "== dist pytest exit $?"; tail -n 4 gpurun_out/pytest_dist.log
for wl in $WLS; do
  timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus $N --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
  echo "== bench $wl N=$N exit $?"; tail -n 3 gpurun_out/bench_${wl}_n$N.err; head -c 300 gpurun_out/bench_${wl}_n$N.json; echo
  timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 --workload $wl --no-cpu-baseline --no-torch-baseline --no-reward > gpurun_out/bench_${wl}_n1_samebox.json 2> gpurun_out/bench_${wl}_n1_samebox.err
  echo "== bench $wl N=1 exit $?"; head -c 300 gpurun_out/bench_${wl}_n1_samebox.json; echo
done
