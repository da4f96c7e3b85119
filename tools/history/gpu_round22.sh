#!/bin/bash
# 8-GPU visit: the fused peer-memory all-reduce at world 8 / 4 / 2 (check + bench), NCCL two-phase at 8 for comparison.
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 8 29611 tools/check_sharded.py --fused > gpurun_out/check_fused_8gpu.json 2>gpurun_out/check_fused_8gpu.err; echo "check_sharded --fused (8) rc=$?"
tail -1 gpurun_out/check_fused_8gpu.json | cut -c1-1500; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/check_fused_8gpu.err | tail -5 | cut -c1-300
summary='
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["n_gpus"], d["config"]["collective"][:44], "value", round(d["value"]/1e9,1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]/1e9,1), round(d["e2e"]["ms_per_step"],4), "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), d["clocks"])'
for n in 8 4 2; do
  run $n 2962$n bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/scale_peer_${n}gpu.log 2>gpurun_out/scale_peer_${n}gpu.err; echo "bench($n, auto) rc=$?"
  python -c "$summary" < gpurun_out/scale_peer_${n}gpu.log; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/scale_peer_${n}gpu.err | tail -3 | cut -c1-300
done
run 8 29631 bench.py --gpus 8 --steps 30 --warmup 5 --collective nccl > gpurun_out/scale_nccl_8gpu.log 2>gpurun_out/scale_nccl_8gpu.err; echo "bench(8, nccl) rc=$?"
python -c "$summary" < gpurun_out/scale_nccl_8gpu.log
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/scale_peer_1gpu.log 2>gpurun_out/scale_peer_1gpu.err; echo "bench(1) rc=$?"
python -c "$summary" < gpurun_out/scale_peer_1gpu.log
