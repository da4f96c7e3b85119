#!/bin/bash
# few-target kernel: two alternating warp sets vs one set (NB200_FT_ONE_SET=1), 2 GiB of X per shape
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "few_target" > gpurun_out/pytest_ft.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ft.log | cut -c1-300
: > gpurun_out/ft_sets.jsonl
for spec in "512 2" "512 3" "512 4" "512 6" "512 10" "256 4" "128 2" "1024 2" "1024 3" "1024 4" "784 3" "1000 4"; do
  set -- $spec; D=$1; T=$2; rows=$((268435456 / D))
  for one in 0 1; do
    NB200_FT_ONE_SET=$one timeout 100 python tools/bench_config3.py --rows $rows --features $D --targets $T --no-cublas --steps 4 > gpurun_out/ft_one.json 2>gpurun_out/ft_one.err || { echo "D=$D T=$T one=$one failed"; tail -1 gpurun_out/ft_one.err; continue; }
    cat gpurun_out/ft_one.json >> gpurun_out/ft_sets.jsonl
    python -c "
import json; d=json.load(open('gpurun_out/ft_one.json')); print($D, $T, d['plan'][:30], 'vgrad', d['vgrad']['ms_med'], 'x_once', d['vgrad']['x_once_gbs'], 'value', d['value']['ms_med'])"
  done
done
