#!/bin/bash
# round 2, GPU call 10 (one GPU): ncu --set full with source for the many-target forward (tensor-map fed and cp.async fed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/bench_config3.py --no-cublas --rows 262144 --steps 2"
$CMD > gpurun_out/r2_10_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mt_forward -s 2 -c 2 -o gpurun_out/r2_10_fwd_tma $CMD > gpurun_out/r2_10_ncu_tma.log 2>&1
echo "ncu tma rc=$?"; tail -3 gpurun_out/r2_10_ncu_tma.log
ls -la gpurun_out/*.ncu-rep
