#!/bin/bash
# One GPU (gpurun -- "bash tools/gpu_ncu.sh"): ncu evidence — launch list of a short bench run, full captures of the hot kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BASE="python bench.py --steps 3 --warmup 3 --no-configs --no-cpu-baseline --sustained-seconds 0 --settle-seconds 0.2"
$BASE > gpurun_out/ncu_plain_main.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_bench.csv $BASE > gpurun_out/ncu_ncu_launches.log 2>&1
echo "launch list rc=$?"
$BASE > gpurun_out/ncu_plain_main.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vgrad_t1 -s 4 -c 1 -o gpurun_out/ncu_t1_2p23x1024 $BASE > gpurun_out/ncu_ncu_t1_main.log 2>&1
echo "t1 main rc=$?"
C2="$BASE --rows-per-gpu 1048576"
$C2 > gpurun_out/ncu_plain_c2.log 2>&1 && \
ncu --set full --clock-control none -k regex:vgrad_t1 -s 4 -c 1 -o gpurun_out/ncu_t1_2p20x1024 $C2 > gpurun_out/ncu_ncu_t1_c2.log 2>&1
echo "t1 cfg2 rc=$?"
C5="$BASE --rows-per-gpu 16777216 --features 512"
$C5 > gpurun_out/ncu_plain_c5.log 2>&1 && \
ncu --set full --clock-control none -k regex:vgrad_t1 -s 4 -c 1 -o gpurun_out/ncu_t1_2p24x512 $C5 > gpurun_out/ncu_ncu_t1_c5.log 2>&1
echo "t1 cfg5 rc=$?"
C3="python tools/bench_config3.py --no-cublas --rows 262144 --steps 2"
$C3 > gpurun_out/ncu_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mt_ -s 3 -c 3 -o gpurun_out/ncu_mt_262144x2048x100 $C3 > gpurun_out/ncu_ncu_mt.log 2>&1
echo "mt rc=$?"
# the reports stay on the box (64 MiB limit on what comes back): export the per-launch metrics, and for the two
# captures with source the per-instruction stall samples; keep only the main kernel's report
for REP in gpurun_out/ncu_*.ncu-rep; do
  ncu -i $REP --page raw --csv > ${REP%.ncu-rep}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/ncu_t1_2p23x1024.ncu-rep --page source --csv > gpurun_out/ncu_t1_2p23x1024_source.csv 2>/dev/null
ncu -i gpurun_out/ncu_mt_262144x2048x100.ncu-rep --page source --csv -k regex:mt_forward --launch-count 1 > gpurun_out/ncu_mt_forward_source.csv 2>/dev/null
ncu -i gpurun_out/ncu_mt_262144x2048x100.ncu-rep --page source --csv -k regex:mt_backward --launch-count 1 > gpurun_out/ncu_mt_backward_source.csv 2>/dev/null
rm -f gpurun_out/ncu_t1_2p20x1024.ncu-rep gpurun_out/ncu_t1_2p24x512.ncu-rep gpurun_out/ncu_mt_262144x2048x100.ncu-rep
ls -la gpurun_out/ | tail -20
