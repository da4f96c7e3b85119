#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_config1.py > gpurun_out/config1.jsonl 2>gpurun_out/config1.err; echo "config1 rc=$?"
cat gpurun_out/config1.jsonl; tail -3 gpurun_out/config1.err
timeout 600 python tools/solve_config.py --config 2 > gpurun_out/solve_config2.jsonl 2>gpurun_out/solve_config2.err; echo "solve2 rc=$?"
cat gpurun_out/solve_config2.jsonl; tail -3 gpurun_out/solve_config2.err
timeout 900 python tools/solve_config.py --config 5 > gpurun_out/solve_config5.jsonl 2>gpurun_out/solve_config5.err; echo "solve5 rc=$?"
cat gpurun_out/solve_config5.jsonl; tail -3 gpurun_out/solve_config5.err
timeout 600 python tools/solve_config.py --config 4 > gpurun_out/solve_config4_1gpu.jsonl 2>gpurun_out/solve_config4.err; echo "solve4(1 gpu, 8Mi rows) rc=$?"
cat gpurun_out/solve_config4_1gpu.jsonl; tail -3 gpurun_out/solve_config4.err
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 240 -o timeout_method=thread -x > gpurun_out/pytest_solvers.log 2>&1; echo "pytest solvers rc=$?"
tail -4 gpurun_out/pytest_solvers.log | cut -c1-300
