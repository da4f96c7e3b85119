#!/bin/bash
# batched models on the one-pass tensor-pipe kernel: parity, then K sweep at D = 1024 and 512
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batch.py -m gpu -q -x --timeout 100 -o timeout_method=thread > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_batch.log | cut -c1-300
timeout 200 python tools/bench_batch.py 1024 > gpurun_out/bench_batch.jsonl 2>gpurun_out/bench_batch.err; echo "batch 1024 rc=$?"; cat gpurun_out/bench_batch.jsonl; tail -2 gpurun_out/bench_batch.err
timeout 200 python tools/bench_batch.py 512 > gpurun_out/bench_batch_512.jsonl 2>gpurun_out/bench_batch.err; echo "batch 512 rc=$?"; cat gpurun_out/bench_batch_512.jsonl
