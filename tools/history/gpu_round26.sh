#!/bin/bash
# new heuristic (rows per iteration / alternating sets): whole GPU suite, bench, sustained sweep with the heuristic choice
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 -o timeout_method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 200 python bench.py --steps 50 --warmup 5 --sustained-seconds 3 > gpurun_out/bench_v6.log 2>gpurun_out/bench_v6.err; echo "bench rc=$?"; cut -c1-2600 gpurun_out/bench_v6.log; tail -2 gpurun_out/bench_v6.err
C="1048576,1024,s-logistic,-1;1048576,1024,mse,-1;2097152,512,s-hinge,-1;4194304,256,s-logistic,-1;8388608,128,s-logistic,-1;16777216,64,s-logistic,-1;1000000,1000,s-logistic,-1;524288,2048,s-logistic,-1;3000000,100,mse,-1"
timeout 300 python tools/bench_sustained.py --seconds 2 --configs "$C" > gpurun_out/sustained_v6.jsonl 2>gpurun_out/sustained_v6.err; echo "sustained rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sustained_v6.jsonl'):
    d=json.loads(l); print(d['rows'], d['features'], d['loss'], d['kernel'][:64], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
tail -3 gpurun_out/sustained_v6.err
