#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 240 -o timeout_method=thread -k "variant or t1" > gpurun_out/pytest_alt.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_alt.log | cut -c1-300
C="${CONFIGS:-1048576,1024,s-logistic,4;1048576,1024,s-logistic,27;1048576,1024,s-logistic,32;1048576,1024,mse,32;2097152,512,s-logistic,33;2097152,512,s-hinge,33;4194304,256,s-logistic,34;16777216,64,s-logistic,36}"
timeout 400 python tools/bench_sustained.py --seconds 2.5 --configs "$C" > gpurun_out/sustained_alt.jsonl 2>gpurun_out/sustained_alt.err; echo "sustained rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/sustained_alt.jsonl'):
    d=json.loads(l); print(d['features'], d['loss'], d['kernel'][:58], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
tail -3 gpurun_out/sustained_alt.err
