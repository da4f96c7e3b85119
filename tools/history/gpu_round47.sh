#!/bin/bash
# end-to-end solves of configs 2, 4 (one 64 GiB shard), 5 with the final kernels
mkdir -p gpurun_out
timeout 300 python tools/solve_config.py --config 2 > gpurun_out/solve_config2.jsonl 2>gpurun_out/solve_config2.err; echo "solve2 rc=$?"; cut -c1-420 gpurun_out/solve_config2.jsonl
timeout 400 python tools/solve_config.py --config 4 > gpurun_out/solve_config4_1gpu.jsonl 2>gpurun_out/solve_config4.err; echo "solve4 rc=$?"; cut -c1-420 gpurun_out/solve_config4_1gpu.jsonl
timeout 600 python tools/solve_config.py --config 5 > gpurun_out/solve_config5.jsonl 2>gpurun_out/solve_config5.err; echo "solve5 rc=$?"; cut -c1-420 gpurun_out/solve_config5.jsonl
