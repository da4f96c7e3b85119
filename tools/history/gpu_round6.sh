#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_solvers.py tests/test_cpp_host.py -m gpu -q --timeout 240 -o timeout_method=thread -x -k "many_target or multi_target or solvers or lbfgs or osga or cpp" > gpurun_out/pytest_mt.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_mt.log | cut -c1-300
timeout 600 python tools/bench_config3.py --rows 1048576 --no-cublas > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 1M rc=$?"
cat gpurun_out/config3_1m.json; tail -3 gpurun_out/config3_1m.err
timeout 900 python tools/bench_config3.py --rows 4194304 --no-cublas > gpurun_out/config3_4m.json 2>gpurun_out/config3_4m.err; echo "config3 4M rc=$?"
cat gpurun_out/config3_4m.json; tail -3 gpurun_out/config3_4m.err
