#!/bin/bash
# round 2, GPU call 1 (one GPU): the new group / concurrency tests, the C++ mirror, then the rebuilt bench.py
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_01_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py -x -q -m gpu > gpurun_out/r2_01_group_tests.log 2>&1
echo "group tests rc=$?" >> gpurun_out/r2_01_group_tests.log
tail -15 gpurun_out/r2_01_group_tests.log
timeout 600 python bench.py > gpurun_out/r2_01_bench_n1.json 2> gpurun_out/r2_01_bench_n1.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2_01_bench_n1.err
head -c 6000 gpurun_out/r2_01_bench_n1.json
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_01_all_tests.log 2>&1
echo "all tests rc=$?" >> gpurun_out/r2_01_all_tests.log
tail -8 gpurun_out/r2_01_all_tests.log
