#!/bin/bash
# round 2, GPU call 14 (TWO GPUs): device groups over real peers, sharded processes under torchrun (fused and NCCL),
# opposite-order evaluation, the C++ mirror on two devices, bench --gpus 2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py tests/test_gpu_solvers.py -x -q -m gpu > gpurun_out/r2_14_group_tests.log 2>&1
echo "group tests rc=$?"; tail -12 gpurun_out/r2_14_group_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_sharded.py --fused > gpurun_out/r2_14_check_fused_2gpu.json 2> gpurun_out/r2_14_check_fused.err
echo "check fused rc=$?"; tail -c 1500 gpurun_out/r2_14_check_fused_2gpu.json; tail -3 gpurun_out/r2_14_check_fused.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 > gpurun_out/r2_14_bench_2gpu.json 2> gpurun_out/r2_14_bench_2gpu.err
echo "bench 2 rc=$?"; tail -c 600 gpurun_out/r2_14_bench_2gpu.err; head -c 1500 gpurun_out/r2_14_bench_2gpu.json
NB200_PEER_WAIT=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --sustained-seconds 0 > gpurun_out/r2_14_bench_2gpu_inkernel.json 2>> gpurun_out/r2_14_bench_2gpu.err
echo "bench 2 in-kernel rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --collective nccl --sustained-seconds 0 > gpurun_out/r2_14_bench_2gpu_nccl.json 2>> gpurun_out/r2_14_bench_2gpu.err
echo "bench 2 nccl rc=$?"
python - <<'PY'
import json
for name in ("bench_2gpu", "bench_2gpu_inkernel", "bench_2gpu_nccl"):
    try:
        d = json.load(open(f"gpurun_out/r2_14_{name}.json"))
        print(name, "ms/step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e9, 1), "e2e ms", round(d["e2e"]["ms_per_step"], 4), "parity", d["parity"]["rel_f"], d["parity"]["rel_g"], d["details"]["collective"][:60], d.get("scaling_limiter"))
    except Exception as error:
        print(name, "unreadable", error)
PY
