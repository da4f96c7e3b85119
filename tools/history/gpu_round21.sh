#!/bin/bash
# lean consumer loop + one-lane loss: parity first, then sustained throughput / power per shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scaling.py -m gpu -q -x --timeout 240 -o timeout_method=thread > gpurun_out/pytest_lean.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_lean.log | cut -c1-300
timeout 300 python tools/bench_sustained.py --seconds 3 > gpurun_out/sustained.jsonl 2>gpurun_out/sustained.err; echo "sustained rc=$?"
cut -c1-420 gpurun_out/sustained.jsonl; tail -3 gpurun_out/sustained.err
timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lean.log 2>gpurun_out/bench_lean.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_lean.log
