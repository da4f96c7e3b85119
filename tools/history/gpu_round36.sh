#!/bin/bash
# tensor-pipe path for any T in [2, 128] and even D: parity on ragged shapes, config 3 unchanged?, a common odd shape
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_solvers.py -m gpu -q -x --timeout 100 -o timeout_method=thread > gpurun_out/pytest_mt.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_mt.log | cut -c1-300
timeout 200 python tools/bench_config3.py --rows 1048576 --no-cublas --steps 6 > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 rc=$?"; cut -c1-600 gpurun_out/config3_1m.json
timeout 200 python tools/bench_config3.py --rows 4194304 --features 784 --targets 10 --no-cublas --steps 6 > gpurun_out/mt_784x10.json 2>gpurun_out/mt_784x10.err; echo "784x10 rc=$?"; cut -c1-600 gpurun_out/mt_784x10.json
timeout 200 python tools/bench_config3.py --rows 4194304 --features 784 --targets 10 --no-cublas --steps 4 --generic > gpurun_out/mt_784x10_generic.json 2>gpurun_out/mt_784x10.err; echo "784x10 generic rc=$?"; cut -c1-600 gpurun_out/mt_784x10_generic.json
