#!/bin/bash
# multi-row reduction + 4-consumer-warp variants: parity, then sustained sweep heuristic vs NCW=4 family
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scaling.py tests/test_gpu_solvers.py -m gpu -q -x --timeout 240 -o timeout_method=thread > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_rows.log | cut -c1-300
C="1048576,1024,s-logistic,-1;1048576,1024,s-logistic,27;1048576,1024,mse,27;2097152,512,s-logistic,-1;2097152,512,s-logistic,28;2097152,512,s-hinge,-1;2097152,512,s-hinge,28;4194304,256,s-logistic,-1;4194304,256,s-logistic,29;8388608,128,s-logistic,-1;8388608,128,s-logistic,30;16777216,64,s-logistic,-1;16777216,64,s-logistic,31"
timeout 400 python tools/bench_sustained.py --seconds 2.5 --configs "$C" > gpurun_out/sustained_rows.jsonl 2>gpurun_out/sustained_rows.err; echo "sustained rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sustained_rows.jsonl'):
    d=json.loads(l); print(d['features'], d['loss'], d['kernel'][:52], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
tail -3 gpurun_out/sustained_rows.err
