#!/bin/bash
# ragged row counts; Y prefetch in the DMMA forward (parity + config 3 at 1 Mi rows); sustained fp64 ceilings
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -m gpu -q -x --timeout 100 -o timeout_method=thread > gpurun_out/pytest_par.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_par.log | cut -c1-300
timeout 200 python tools/bench_config3.py --rows 1048576 --no-cublas --steps 8 > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 rc=$?"; cut -c1-700 gpurun_out/config3_1m.json
timeout 100 tools/bin/fp64_peaks --sustained > gpurun_out/fp64_peaks.jsonl 2>gpurun_out/fp64_peaks.err; echo "peaks rc=$?"; cat gpurun_out/fp64_peaks.jsonl
timeout 300 python tools/bench_config3.py --rows 4194304 --no-cublas --steps 20 > gpurun_out/config3_4m.json 2>gpurun_out/config3_4m.err; echo "config3 4M rc=$?"; cut -c1-700 gpurun_out/config3_4m.json
