#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/profile_layer.py"
$CMD > gpurun_out/layer_plain.log 2>&1 && cat gpurun_out/layer_plain.log &&
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm|tc_wgrad" -s 6 -c 6 -f -o gpurun_out/prof_layer $CMD > gpurun_out/ncu_layer.log 2>&1
echo "== full capture exit $?"
ncu -i gpurun_out/prof_layer.ncu-rep --page raw --csv > gpurun_out/prof_layer_raw.csv 2> /dev/null
ncu -i gpurun_out/prof_layer.ncu-rep --page source --csv --kernel-name regex:tc_wgrad > gpurun_out/prof_layer_wgrad_source.csv 2> /dev/null
ls -la gpurun_out/ | grep prof_layer
sz=$(stat -c %s gpurun_out/prof_layer.ncu-rep 2>/dev/null || echo 0)
if [ "$sz" -gt 30000000 ]; then rm -f gpurun_out/prof_layer.ncu-rep; fi
