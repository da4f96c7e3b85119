#!/bin/bash
# One GPU: ncu --set full of the many-target pair (config 3's width, 2^18 rows), raw metrics + top stall lines
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
C3="python tools/bench_config3.py --no-cublas --rows 262144 --steps 2"
timeout 120 $C3 > gpurun_out/ncu_plain_c3.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mt_ -s 3 -c 3 -o gpurun_out/ncu_mt_262144x2048x100 $C3 > gpurun_out/ncu_ncu_mt.log 2>&1
echo "mt rc=$?"
ncu -i gpurun_out/ncu_mt_262144x2048x100.ncu-rep --page raw --csv > gpurun_out/ncu_mt_262144x2048x100_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_mt_262144x2048x100.ncu-rep --page source --csv -k regex:mt_forward --launch-count 1 > gpurun_out/ncu_mt_forward_source.csv 2>/dev/null
ncu -i gpurun_out/ncu_mt_262144x2048x100.ncu-rep --page source --csv -k regex:mt_backward --launch-count 1 > gpurun_out/ncu_mt_backward_source.csv 2>/dev/null
rm -f gpurun_out/ncu_mt_262144x2048x100.ncu-rep
ls -la gpurun_out | grep ncu_mt
