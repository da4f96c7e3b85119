#!/bin/bash
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
run 2 29811 tools/check_sharded.py --fused > gpurun_out/check_fused_2gpu.json 2>gpurun_out/check_fused_2gpu.err; echo "check_sharded --fused (2) rc=$?"
tail -1 gpurun_out/check_fused_2gpu.json | cut -c1-1300; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/check_fused_2gpu.err | tail -3 | cut -c1-300
