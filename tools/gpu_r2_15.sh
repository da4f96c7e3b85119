#!/bin/bash
# round 2, GPU call 15 (EIGHT GPUs): sharded processes (fused peers, opposite order), an 8-device group in one process
# (Python and the C++ mirror), bench --gpus 8 (2^26 global rows)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tools/check_sharded.py --fused > gpurun_out/r2_15_check_fused_8gpu.json 2> gpurun_out/r2_15_check_fused.err
echo "check fused rc=$?"; tail -c 700 gpurun_out/r2_15_check_fused_8gpu.json; tail -2 gpurun_out/r2_15_check_fused.err
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_cpp_host.py -x -q -m gpu -k "not torchrun" > gpurun_out/r2_15_group_tests.log 2>&1
echo "group tests rc=$?"; tail -4 gpurun_out/r2_15_group_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 > gpurun_out/r2_15_bench_8gpu.json 2> gpurun_out/r2_15_bench_8gpu.err
echo "bench 8 rc=$?"; tail -c 400 gpurun_out/r2_15_bench_8gpu.err
timeout 300 python tools/bench_group.py --devices 8 > gpurun_out/r2_15_group_8dev.json 2> gpurun_out/r2_15_group.err
echo "group bench rc=$?"; cat gpurun_out/r2_15_group_8dev.json; tail -3 gpurun_out/r2_15_group.err
python - <<'PY'
import json
lines = [l for l in open("gpurun_out/r2_15_bench_8gpu.json").read().split("\n") if l.startswith("{")]
d = json.loads(lines[-1])
print("8 GPUs: ms/step", d["ms_per_step"], "value", d["value"] / 1e9, "global rows", d["config"]["global_rows"], "parity", d["parity"], d.get("scaling_limiter"))
for r in d["per_rank"]:
    print(" rank", r["rank"], round(r["kernel_ms"]["median"], 4), round(r["step_ms"], 4), r["clocks"]["sm_mhz"], r["clocks"]["power_w"])
print("sustained", d.get("sustained"))
PY
