#!/bin/bash
# T in {2, 3, 4}: CUDA-core one-pass (-6) vs tensor-pipe one-pass with two warp sets (-5)
mkdir -p gpurun_out
: > gpurun_out/fd_small_t.jsonl
for spec in "512 2" "512 3" "512 4" "768 2" "768 3" "768 4" "1024 2" "1024 3" "1024 4" "256 4"; do
  set -- $spec; D=$1; T=$2; rows=$((268435456 / D))
  for v in -5 -6; do
    timeout 100 python tools/bench_config3.py --rows $rows --features $D --targets $T --no-cublas --steps 4 --variant $v > gpurun_out/fd_one.json 2>gpurun_out/fd_one.err || { echo "D=$D T=$T v=$v failed: $(tail -1 gpurun_out/fd_one.err | cut -c1-150)"; continue; }
    cat gpurun_out/fd_one.json >> gpurun_out/fd_small_t.jsonl
    python -c "
import json; d=json.load(open('gpurun_out/fd_one.json')); print($D, $T, d['plan'][:40], 'vgrad', d['vgrad']['ms_med'], 'x_once', d['vgrad']['x_once_gbs'], 'value', d['value']['ms_med'])"
  done
done
