#!/bin/bash
# one-pass tensor-pipe kernel: two alternating warp sets (T <= 8) vs one set (NB200_FD_ONE_SET=1)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "handful" > gpurun_out/pytest_fd.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_fd.log | cut -c1-300
: > gpurun_out/fd_sets.jsonl
for spec in "1024 8" "1000 5" "1024 6" "768 8" "512 6" "512 8" "300 7"; do
  set -- $spec; D=$1; T=$2; rows=$((268435456 / D))
  for one in 0 1; do
    NB200_FD_ONE_SET=$one timeout 100 python tools/bench_config3.py --rows $rows --features $D --targets $T --no-cublas --steps 4 --variant -5 > gpurun_out/fd_one.json 2>gpurun_out/fd_one.err || { echo "D=$D T=$T one=$one failed: $(tail -1 gpurun_out/fd_one.err | cut -c1-150)"; continue; }
    cat gpurun_out/fd_one.json >> gpurun_out/fd_sets.jsonl
    python -c "
import json; d=json.load(open('gpurun_out/fd_one.json')); print($D, $T, d['plan'][:40], 'vgrad', d['vgrad']['ms_med'], 'x_once', d['vgrad']['x_once_gbs'], 'value', d['value']['ms_med'])"
  done
done
