#!/bin/bash
# rebuilt library (loss error / evaluate kernels): whole GPU suite incl. the new error, evaluate and linear-model tests,
# then a short bench to confirm the hot kernel is unchanged
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q -x --timeout 120 -o timeout_method=thread > gpurun_out/pytest_gpu_r52.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r52.log
tail -4 gpurun_out/pytest_gpu_r52.log
timeout 40 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r52.log 2> gpurun_out/bench_r52.err
echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_r52.log
