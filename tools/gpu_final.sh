#!/bin/bash
# One GPU, the short form of gpu_suite.sh: the whole GPU suite and the bench line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --maxfail=10 > gpurun_out/all_tests.log 2>&1
echo "all tests rc=$?"; tail -3 gpurun_out/all_tests.log
timeout 240 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.err
