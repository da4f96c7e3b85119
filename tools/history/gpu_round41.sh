#!/bin/bash
# one-pass tensor-pipe kernel for a handful of targets: parity, then T / D sweep against the other paths
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "handful or few_target" > gpurun_out/pytest_fd.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_fd.log | cut -c1-300
: > gpurun_out/fd_sweep.jsonl
for spec in "784 10" "1000 5" "1024 8" "512 16" "512 10" "512 6" "512 4" "1024 4" "256 10" "100 10"; do
  set -- $spec; D=$1; T=$2; rows=$((268435456 / D))
  for v in -5 -6 -4; do
    timeout 100 python tools/bench_config3.py --rows $rows --features $D --targets $T --no-cublas --steps 4 --variant $v > gpurun_out/fd_one.json 2>gpurun_out/fd_one.err || { echo "D=$D T=$T v=$v failed: $(tail -1 gpurun_out/fd_one.err | cut -c1-150)"; continue; }
    cat gpurun_out/fd_one.json >> gpurun_out/fd_sweep.jsonl
    python -c "
import json; d=json.load(open('gpurun_out/fd_one.json')); print($D, $T, d['plan'][:34], 'vgrad', d['vgrad']['ms_med'], 'x_once', d['vgrad']['x_once_gbs'], 'value', d['value']['ms_med'])"
  done
done
