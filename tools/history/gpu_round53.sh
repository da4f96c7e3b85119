#!/bin/bash
# app/bench_linear.cpp over the GPU path: nested cross-validation of ridge + mse (device-resident L-BFGS) and of
# lasso + mae (host-driven RQB), plus the throughput of linear::evaluate on its own
mkdir -p gpurun_out
timeout 60 python tools/bench_linear.py --linear ridge --loss mse --rows 262144 --features 256 > gpurun_out/bench_linear_ridge.jsonl 2> gpurun_out/bench_linear_ridge.err
echo "ridge rc=$?"
timeout 70 python tools/bench_linear.py --linear lasso --loss mae --rows 32768 --features 32 --tuner-max-evals 4 > gpurun_out/bench_linear_lasso.jsonl 2> gpurun_out/bench_linear_lasso.err
echo "lasso rc=$?"
cut -c1-600 gpurun_out/bench_linear_ridge.jsonl gpurun_out/bench_linear_lasso.jsonl
tail -3 gpurun_out/bench_linear_ridge.err gpurun_out/bench_linear_lasso.err
