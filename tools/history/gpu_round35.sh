#!/bin/bash
# 2-GPU re-check of the fused peer path with the current kernels (variant 27 / 37), sharded CPU-vs-GPU parity, bench
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 2 29711 tools/check_sharded.py --fused > gpurun_out/check_fused_2gpu.json 2>gpurun_out/check_fused_2gpu.err; echo "check_sharded --fused (2) rc=$?"
tail -1 gpurun_out/check_fused_2gpu.json | cut -c1-1200; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/check_fused_2gpu.err | tail -4 | cut -c1-300
run 2 29712 tools/check_sharded.py > gpurun_out/check_sharded_2gpu.json 2>gpurun_out/check_sharded_2gpu.err; echo "check_sharded nccl (2) rc=$?"
tail -1 gpurun_out/check_sharded_2gpu.json | cut -c1-400
summary='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["config"]["collective"][:44], "value", round(d["value"]/1e9,1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e9,1), round(d["e2e"]["ms_per_step"],4), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), d["config"]["kernel"][:40])'
run 2 29713 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_2gpu.log 2>gpurun_out/bench_2gpu.err; echo "bench(2) rc=$?"; python -c "$summary" < gpurun_out/bench_2gpu.log
run 2 29714 bench.py --gpus 2 --steps 20 --warmup 3 --rows-per-gpu 8388608 > gpurun_out/bench_2gpu_config4.log 2>gpurun_out/bench_2gpu_config4.err; echo "bench(2, 2^23 rows) rc=$?"; python -c "$summary" < gpurun_out/bench_2gpu_config4.log
