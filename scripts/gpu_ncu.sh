#!/bin/bash
# ncu evidence for profiles/ (usage: bash scripts/gpu_ncu.sh): launch list of a short bench run, then
# `--set full` captures of every weight-gradient launch of ONE step (the bench line's roofline family),
# of the recurrent kernels and of a few persistent conv GEMMs.  VAR_GRU_NO_COOP=1: ncu cannot replay the
# cooperative cluster launch of the BPTT kernel.
mkdir -p gpurun_out
export VAR_GRU_NO_COOP=1
CMD="python scripts/profile_step.py"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "== launch list exit $?"
# 3 warm-up steps x 16 weight-gradient launches are skipped: launches 48..63 are the measured step
ncu --set full --clock-control none --import-source on -k regex:"wgrad" -s 48 -c 16 -f -o gpurun_out/prof_wgrad $CMD > gpurun_out/ncu_full_wgrad.log 2>&1
echo "== wgrad capture exit $?"
ncu -i gpurun_out/prof_wgrad.ncu-rep --page raw --csv > gpurun_out/prof_wgrad_raw.csv 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:"gru_persist|gru_bwd_ksplit|mfcc" -s 9 -c 3 -f -o gpurun_out/prof_gru $CMD > gpurun_out/ncu_full_gru.log 2>&1
echo "== gru/mfcc capture exit $?"
ncu -i gpurun_out/prof_gru.ncu-rep --page raw --csv > gpurun_out/prof_gru_raw.csv 2> /dev/null
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_persist" -s 66 -c 22 -f -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full_gemm.log 2>&1
echo "== gemm capture exit $?"
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2> /dev/null
# the im2col-free conv kernels (6 launches per step: snd.conv2/3 forward + data gradient, imgBranch.2/.5 forward)
ncu --set full --clock-control none --import-source on -k regex:"halo_conv" -s 18 -c 6 -f -o gpurun_out/prof_halo $CMD > gpurun_out/ncu_full_halo.log 2>&1
echo "== halo capture exit $?"
ncu -i gpurun_out/prof_halo.ncu-rep --page raw --csv > gpurun_out/prof_halo_raw.csv 2> /dev/null
for f in prof_wgrad prof_gru prof_gemm prof_halo; do
  sz=$(stat -c %s gpurun_out/$f.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 14000000 ]; then echo "dropping oversized report $f ($sz bytes)"; rm -f gpurun_out/$f.ncu-rep; fi
done
ls -la gpurun_out/ | grep -E "prof_|launches"
