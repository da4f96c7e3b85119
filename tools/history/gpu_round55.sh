#!/bin/bash
# library with the host-side bundle QP (nb200_simplex_qp): solver + linear-model GPU tests, bench_linear lasso + mae
# (host-driven RQB), short bench
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_solvers.py tests/test_linear_model.py tests/test_loss_error.py -m gpu -q -x --timeout 30 -o timeout_method=thread > gpurun_out/pytest_gpu_r55.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r55.log
tail -n 3 gpurun_out/pytest_gpu_r55.log
timeout 30 python tools/bench_linear.py --linear lasso --loss mae --rows 32768 --features 32 --tuner-max-evals 4 > gpurun_out/bench_linear_lasso_r55.jsonl 2> gpurun_out/bench_linear_lasso_r55.err
echo "lasso rc=$?"
cut -c1-420 gpurun_out/bench_linear_lasso_r55.jsonl
timeout 25 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r55.log 2> gpurun_out/bench_r55.err
echo "bench rc=$?"
cut -c1-200 gpurun_out/bench_r55.log
