#!/bin/bash
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log | cut -c1-260
timeout 200 python -m pytest tests/test_cpp_host.py tests/test_gpu_solvers.py -m gpu -q -x --timeout 100 -o timeout_method=thread > gpurun_out/pytest_cpp.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_cpp.log | cut -c1-300
