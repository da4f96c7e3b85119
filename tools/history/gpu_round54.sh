#!/bin/bash
# nb200_evaluate on a grow-only scratch arena + pinned staging: the evaluate / linear-model tests, bench_linear (ridge),
# smoke, and a short bench (hot kernel unchanged)
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_loss_error.py tests/test_linear_model.py tests/test_gpu_parity.py -m gpu -q -x --timeout 50 -o timeout_method=thread > gpurun_out/pytest_gpu_r54.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r54.log
tail -n 3 gpurun_out/pytest_gpu_r54.log
timeout 40 python tools/bench_linear.py --linear ridge --loss mse --rows 262144 --features 256 > gpurun_out/bench_linear_ridge_r54.jsonl 2> gpurun_out/bench_linear_ridge_r54.err
echo "ridge rc=$?"
cut -c1-330 gpurun_out/bench_linear_ridge_r54.jsonl
timeout 30 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r54.log 2> gpurun_out/bench_r54.err
echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_r54.log
