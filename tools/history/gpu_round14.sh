#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -q --timeout 240 -o timeout_method=thread -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest.log | cut -c1-300
timeout 900 python tools/bench_config3.py --rows 4194304 --no-cublas > gpurun_out/config3_4m.json 2>gpurun_out/config3_4m.err; echo "config3 4M rc=$?"
cat gpurun_out/config3_4m.json; tail -3 gpurun_out/config3_4m.err
timeout 600 python tools/bench_batch.py > gpurun_out/bench_batch.jsonl 2>gpurun_out/bench_batch.err; echo "bench_batch rc=$?"
cat gpurun_out/bench_batch.jsonl; tail -3 gpurun_out/bench_batch.err
