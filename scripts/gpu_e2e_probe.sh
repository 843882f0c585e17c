#!/bin/bash
# Where does the e2e gap come from?  (diagnostic switches of bench.py; not reported lines)
mkdir -p gpurun_out
for cfg in "A:" "B:VAR_E2E_NOSYNC=1"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs python bench.py --workload ithor_b256 --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-reward 2> gpurun_out/e2e_$name.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$name [$envs] value %.3f ms  e2e %.3f ms  lagged %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pipelined_read']['ms_per_step']), d['e2e'].get('producer_ms_per_batch'))"
done
