#!/bin/bash
# config-4 shard size (2^23 rows x 1024 = 64 GiB per GPU): whole-rows variant 27 vs row-split variant 37
mkdir -p gpurun_out
for v in 27 37 27 37; do
  timeout 300 python bench.py --rows-per-gpu 8388608 --steps 30 --warmup 3 --variant $v --no-cpu-baseline --settle-seconds 3 > gpurun_out/c4_$v.log 2>gpurun_out/c4.err; echo "variant $v rc=$?"
  python - $v <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/c4_{sys.argv[1]}.log').read().strip().splitlines()[-1])
print('value', round(d['value']/1e9,1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e9,1), round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks'])
PY
done
