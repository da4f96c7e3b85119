#!/bin/bash
# One GPU: the many-target pair after an epilogue change — parity of the tensor-pipe paths, then config 3 twice
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -x -q -m gpu -k "many_target or multi_target or config3 or golden or extreme or large_output" > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/c3_tests.log
: > gpurun_out/c3_ab.jsonl
for i in 1 2; do
  timeout 200 python tools/bench_config3.py --no-cublas --steps 5 >> gpurun_out/c3_ab.jsonl 2>> gpurun_out/c3.err
done
cut -c1-400 gpurun_out/c3_ab.jsonl
