#!/bin/bash
# few-target kernel vs tensor-pipe pair over T and D (2^31 bytes of X each): where is the crossover?
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "few_target" > gpurun_out/pytest_ft.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ft.log | cut -c1-300
: > gpurun_out/ft_sweep.jsonl
for D in 512 1024; do
  rows=$((268435456 / D))
  for T in 2 3 4 5 6 8 10; do
    for v in -1 -4; do
      timeout 100 python tools/bench_config3.py --rows $rows --features $D --targets $T --no-cublas --steps 4 --variant $v > gpurun_out/ft_one.json 2>gpurun_out/ft_one.err || { echo "D=$D T=$T v=$v failed"; tail -1 gpurun_out/ft_one.err; continue; }
      cat gpurun_out/ft_one.json >> gpurun_out/ft_sweep.jsonl
      python -c "
import json; d=json.load(open('gpurun_out/ft_one.json')); print($D, $T, d['plan'][:24], 'vgrad', d['vgrad']['ms_med'], 'x_once', d['vgrad']['x_once_gbs'], 'value', d['value']['ms_med'])"
    done
  done
done
