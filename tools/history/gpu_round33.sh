#!/bin/bash
# multi-TU build: whole GPU suite, bench, sustained sweep with the heuristic, config 3
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 120 -o timeout_method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python bench.py --steps 50 --warmup 5 --sustained-seconds 3 --no-cpu-baseline > gpurun_out/bench_v7.log 2>gpurun_out/bench_v7.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v7.log').read().strip().splitlines()[-1])
print('value', round(d['value']/1e9,1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']/1e9,1), round(d['e2e']['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'sustained', d.get('sustained'), d['config']['kernel'])
PY
C="1048576,1024,s-logistic,-1;1048576,1024,s-logistic,37;1048576,1024,mse,-1;2097152,512,s-hinge,-1;4194304,256,s-logistic,-1;8388608,128,s-logistic,-1;16777216,64,s-logistic,-1;1000000,1000,s-logistic,-1;524288,2048,s-logistic,-1"
timeout 300 python tools/bench_sustained.py --seconds 2 --configs "$C" > gpurun_out/sustained_v7.jsonl 2>gpurun_out/sustained_v7.err; echo "sustained rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sustained_v7.jsonl'):
    d=json.loads(l); print(d['rows'], d['features'], d['loss'], d['kernel'][:64], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
timeout 200 python tools/bench_config3.py --rows 1048576 --no-cublas --steps 6 > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 rc=$?"; cut -c1-600 gpurun_out/config3_1m.json
