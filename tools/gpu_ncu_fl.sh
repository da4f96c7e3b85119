#!/bin/bash
# One GPU: ncu --set full of the lane-loss kernels (vgrad_fl.cuh) at 784 x 10 (pipeline) and 512 x 16 (two barriers)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
A="python tools/sweep_multi.py --shapes 784,10,s-classnll --gib 0.5 --seconds 0.1"
timeout 120 $A > gpurun_out/ncu_plain_fp.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vgrad_fp -s 3 -c 1 -o gpurun_out/ncu_fp_784x10 $A > gpurun_out/ncu_ncu_fp.log 2>&1
echo "fp rc=$?"
B="python tools/sweep_multi.py --shapes 512,16,s-classnll --gib 0.5 --seconds 0.1"
timeout 120 $B > gpurun_out/ncu_plain_fl.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vgrad_fl -s 3 -c 1 -o gpurun_out/ncu_fl_512x16 $B > gpurun_out/ncu_ncu_fl.log 2>&1
echo "fl rc=$?"
for NAME in fp_784x10 fl_512x16; do
  ncu -i gpurun_out/ncu_$NAME.ncu-rep --page raw --csv > gpurun_out/ncu_${NAME}_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$NAME.ncu-rep --page source --csv > gpurun_out/ncu_${NAME}_source.csv 2>/dev/null
done
rm -f gpurun_out/ncu_fl_512x16.ncu-rep
ls -la gpurun_out/ | grep ncu_f
