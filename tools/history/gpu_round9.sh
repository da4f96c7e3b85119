#!/bin/bash
# 2-GPU visit: fused peer-memory all-reduce.
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/check_sharded.py --fused > gpurun_out/check_fused_2gpu.json 2>gpurun_out/check_fused_2gpu.err; echo "check_sharded --fused (2) rc=$?"
tail -1 gpurun_out/check_fused_2gpu.json | cut -c1-1500; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/check_fused_2gpu.err | tail -8 | cut -c1-400
for mode in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2956$((RANDOM % 9)) bench.py --gpus 2 --steps 50 --warmup 5 --collective $mode > gpurun_out/bench_2gpu_$mode.log 2>gpurun_out/bench_2gpu_$mode.err; echo "bench(2, $mode) rc=$?"
  tail -1 gpurun_out/bench_2gpu_$mode.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['config']['collective'][:40], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
  grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_2gpu_$mode.err | tail -4 | cut -c1-300
done
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>&1; tail -1 gpurun_out/bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])"
