#!/bin/bash
# A/B on one box: Y prefetch in the DMMA forward epilogue
mkdir -p gpurun_out
for pf in 0 1 0 1; do
  NB200_MT_PREFETCH_Y=$pf timeout 200 python tools/bench_config3.py --rows 1048576 --no-cublas --steps 8 > gpurun_out/c3_$pf.json 2>gpurun_out/c3.err; echo "prefetch=$pf rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_$pf.json')); print(d['vgrad'], d['value'])"
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
