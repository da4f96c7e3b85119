#!/bin/bash
# 8 x B200 at the config-4 shard size (2^23 rows = 64 GiB per GPU, 64 Mi rows in total): weak-scaling bench + the solve
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
summary='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["config"]["collective"][:44], "value", round(d["value"]/1e9,1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e9,1), round(d["e2e"]["ms_per_step"],4), "frac", round(d["roofline"]["frac"],3), d["config"]["kernel"][:40], d["clocks"])'
run 8 29911 bench.py --gpus 8 --steps 20 --warmup 3 --rows-per-gpu 8388608 > gpurun_out/bench_8gpu_config4.log 2>gpurun_out/bench_8gpu_config4.err; echo "bench(8, 2^23 rows) rc=$?"; python -c "$summary" < gpurun_out/bench_8gpu_config4.log; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_8gpu_config4.err | tail -2 | cut -c1-300
run 8 29912 tools/solve_config.py --config 4 > gpurun_out/solve_config4_8gpu.jsonl 2>gpurun_out/solve_config4_8gpu.err; echo "solve4(8) rc=$?"; cut -c1-520 gpurun_out/solve_config4_8gpu.jsonl
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 3 --rows-per-gpu 8388608 --no-cpu-baseline > gpurun_out/bench_1gpu_config4.log 2>gpurun_out/bench_1gpu_config4.err; echo "bench(1, 2^23 rows) rc=$?"; python -c "$summary" < gpurun_out/bench_1gpu_config4.log
