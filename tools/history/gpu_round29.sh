#!/bin/bash
# rows split over two warps with several rows per iteration (variants 37-39) for D = 1024 / 2048
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 60 -o timeout_method=thread -k "variant" > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ws.log | cut -c1-300
C="1048576,1024,s-logistic,27;1048576,1024,s-logistic,37;1048576,1024,s-logistic,38;1048576,1024,mse,37;524288,2048,s-logistic,5;524288,2048,s-logistic,39;1000000,1000,s-logistic,37"
timeout 300 python tools/bench_sustained.py --seconds 2.5 --configs "$C" > gpurun_out/sustained_ws.jsonl 2>gpurun_out/sustained_ws.err; echo "sustained rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sustained_ws.jsonl'):
    d=json.loads(l); print(d['rows'], d['features'], d['loss'], d['kernel'][:58], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
tail -3 gpurun_out/sustained_ws.err
