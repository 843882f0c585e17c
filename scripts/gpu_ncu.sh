#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + a small full capture of the
# GEMM kernels.  Keeps gpurun_out/ well under the 64 MiB pull limit.
mkdir -p gpurun_out
CMD="python scripts/profile_step.py"
SKIP=${SKIP:-1100}
COUNT=${COUNT:-8}
KRE=${KRE:-tc_gemm|tc_wgrad}
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "== launch list exit $?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c $COUNT -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "== full capture exit $?"
ncu -i gpurun_out/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2> /dev/null
ls -la gpurun_out/
sz=$(stat -c %s gpurun_out/prof.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 45000000 ]; then echo "dropping oversized report ($sz bytes)"; rm -f gpurun_out/prof.ncu-rep; fi
