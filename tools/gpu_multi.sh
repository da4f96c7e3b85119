#!/bin/bash
# Multi-GPU validation on N GPUs of one box (gpurun --gpus N -- 'bash tools/gpu_multi.sh N'): sharded processes under
# torchrun (peers + NCCL, opposite-order evaluation), device groups in one process (Python and the C++ mirror), the
# bench line at N GPUs and the single-process group beside one device.
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    tools/check_sharded.py --fused > gpurun_out/check_fused_${N}gpu.json 2> gpurun_out/check_fused_${N}gpu.err
echo "check fused rc=$?"; tail -c 400 gpurun_out/check_fused_${N}gpu.json
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py tests/test_gpu_solvers.py -x -q -m gpu \
    > gpurun_out/group_tests_${N}gpu.log 2>&1
echo "group tests rc=$?"; tail -3 gpurun_out/group_tests_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 \
    bench.py --gpus $N > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${N}gpu.err
timeout 300 python tools/bench_group.py --devices $N > gpurun_out/group_${N}dev.json 2> gpurun_out/group_${N}dev.err
echo "group bench rc=$?"; cat gpurun_out/group_${N}dev.json
