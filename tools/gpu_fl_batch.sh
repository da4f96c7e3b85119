#!/bin/bash
# One GPU: the batched-models path on the lane-loss kernels — parity, then K models per pass at 2^20 x 1024 / 512 (A/B)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py tests/test_gpu_batch_trials.py -x -q -m gpu > gpurun_out/fl_batch_tests.log 2>&1
echo "batch tests rc=$?"; tail -3 gpurun_out/fl_batch_tests.log
: > gpurun_out/fl_batch.jsonl
for fl in 0 1; do
  for D in 1024 512; do
    echo "{\"NB200_FL\": $fl, \"D\": $D}" >> gpurun_out/fl_batch.jsonl
    NB200_FL=$fl timeout 120 python tools/bench_batch.py $D >> gpurun_out/fl_batch.jsonl 2>> gpurun_out/fl_batch.err
  done
done
cut -c1-160 gpurun_out/fl_batch.jsonl
