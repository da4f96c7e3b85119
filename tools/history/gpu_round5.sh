#!/bin/bash
# Fifth GPU visit (2 GPUs): NCCL path — sharded check, bench at N=2, and the new device-resident L-BFGS tests.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_gpu_solvers.py -m gpu -q --timeout 240 -o timeout_method=thread -x > gpurun_out/pytest_solvers.log 2>&1; echo "pytest solvers rc=$?"
tail -8 gpurun_out/pytest_solvers.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/check_sharded_2gpu.json 2>gpurun_out/check_sharded_2gpu.err; echo "check_sharded(2) rc=$?"
tail -2 gpurun_out/check_sharded_2gpu.json | cut -c1-1200; tail -5 gpurun_out/check_sharded_2gpu.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2>gpurun_out/bench_2gpu.err; echo "bench(2) rc=$?"
tail -1 gpurun_out/bench_2gpu.log | cut -c1-1600; tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1gpu.log 2>gpurun_out/bench_1gpu.err; echo "bench(1) rc=$?"
tail -1 gpurun_out/bench_1gpu.log | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --rows-per-gpu 8388608 > gpurun_out/bench_2gpu_config4.log 2>gpurun_out/bench_2gpu_config4.err; echo "bench(2, 8Mi rows/gpu) rc=$?"
tail -1 gpurun_out/bench_2gpu_config4.log | cut -c1-1600
