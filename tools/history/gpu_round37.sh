#!/bin/bash
# few-target one-pass kernel: parity, then the 4 Mi x 784 x 10 softmax problem on the three paths
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "few_target or many_target or multi_target" > gpurun_out/pytest_ft.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_ft.log | cut -c1-400
for v in "" "--variant -4"; do
timeout 200 python tools/bench_config3.py --rows 4194304 --features 784 --targets 10 --no-cublas --steps 6 $v > gpurun_out/ft_784x10.json 2>gpurun_out/ft_784x10.err; echo "784x10 $v rc=$?"; cut -c1-600 gpurun_out/ft_784x10.json; tail -2 gpurun_out/ft_784x10.err
done
timeout 200 python tools/bench_config3.py --rows 8388608 --features 512 --targets 2 --no-cublas --steps 6 > gpurun_out/ft_512x2.json 2>gpurun_out/ft_512x2.err; echo "512x2 rc=$?"; cut -c1-600 gpurun_out/ft_512x2.json
timeout 200 python tools/bench_config3.py --rows 4194304 --features 1024 --targets 4 --no-cublas --steps 6 > gpurun_out/ft_1024x4.json 2>gpurun_out/ft_1024x4.err; echo "1024x4 rc=$?"; cut -c1-600 gpurun_out/ft_1024x4.json
