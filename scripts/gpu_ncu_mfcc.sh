#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_mfcc.py > gpurun_out/mfcc_plain.log 2>&1; cat gpurun_out/mfcc_plain.log
CMD="python scripts/profile_mfcc.py 0"
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mfcc_kernel -s 3 -c 1 -f -o gpurun_out/prof_mfcc $CMD > gpurun_out/ncu_mfcc.log 2>&1
echo "== capture exit $?"
ncu -i gpurun_out/prof_mfcc.ncu-rep --page raw --csv > gpurun_out/prof_mfcc_raw.csv 2> /dev/null
ncu -i gpurun_out/prof_mfcc.ncu-rep --page source --csv > gpurun_out/prof_mfcc_source.csv 2> /dev/null
ls -la gpurun_out | grep mfcc
