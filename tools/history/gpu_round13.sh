#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -q --timeout 240 -o timeout_method=thread -x -k "batch or many_target or multi_target" > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_batch.log | cut -c1-300
timeout 600 python tools/bench_batch.py > gpurun_out/bench_batch.jsonl 2>gpurun_out/bench_batch.err; echo "bench_batch rc=$?"
cat gpurun_out/bench_batch.jsonl; tail -3 gpurun_out/bench_batch.err
