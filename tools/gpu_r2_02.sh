#!/bin/bash
# round 2, GPU call 2 (one GPU): D = 1024 variant sweep, burst + steady, baseline logistic chain vs the select form
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CFG="1048576,1024,s-logistic,27;1048576,1024,s-logistic,32;1048576,1024,s-logistic,37;1048576,1024,s-logistic,38;1048576,1024,mse,27;1048576,1024,mse,37;1048576,1024,s-hinge,27;2097152,512,s-logistic,33;2097152,512,s-logistic,28"
NB200_LIBRARY=$PWD/tools/bin/libnb200_base.so timeout 600 python tools/bench_sustained.py --seconds 2.5 --configs "$CFG" > gpurun_out/r2_02_sweep_base.jsonl 2> gpurun_out/r2_02_sweep_base.err
echo "base rc=$?"
timeout 600 python tools/bench_sustained.py --seconds 2.5 --configs "$CFG" > gpurun_out/r2_02_sweep_select.jsonl 2> gpurun_out/r2_02_sweep_select.err
echo "select rc=$?"
cat gpurun_out/r2_02_sweep_base.jsonl gpurun_out/r2_02_sweep_select.jsonl | cut -c1-100
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -x -q -m gpu 2>&1 | tail -3
