#!/bin/bash
# One GPU: the row-split / two-set kernels after a loss-step change — their parity tests, then the short-row sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -m gpu -k "short_rows or handful or feature_split or few_target or large_output or extreme or lane_loss or many_targets or repeated" > gpurun_out/fr_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/fr_tests.log
timeout 200 python tools/sweep_multi.py --shapes "256,10,s-classnll;256,16,s-classnll;100,10,s-classnll;64,10,s-classnll;200,5,s-classnll;128,4,s-classnll;256,2,s-classnll;768,8,s-classnll;512,16,s-classnll;384,10,s-classnll;512,8,s-classnll" --gib 2 --seconds 0.4 > gpurun_out/fr_sweep.jsonl 2> gpurun_out/fr_sweep.err
cut -c1-60,150-330 gpurun_out/fr_sweep.jsonl
