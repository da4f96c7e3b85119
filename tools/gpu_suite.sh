#!/bin/bash
# One GPU (gpurun -- "bash tools/gpu_suite.sh"): the whole GPU suite, both bench arms, the batch bench, the lane-loss sweep.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/all_tests.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/all_tests.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.err
timeout 200 python tools/bench_batch.py > gpurun_out/bench_batch.jsonl 2>&1; tail -4 gpurun_out/bench_batch.jsonl
FL_SKIP_TESTS=1 bash tools/gpu_fl.sh
timeout 300 python bench.py --impl reference --steps 16 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference rc=$?"; head -c 400 gpurun_out/bench_reference.json
