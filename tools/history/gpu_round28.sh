#!/bin/bash
# device-resident CGD tests; suspend-time hint on the consumers' stage wait (power / steady-state effect)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_solvers.py -m gpu -q -x --timeout 100 -o timeout_method=thread > gpurun_out/pytest_cgd.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_cgd.log | cut -c1-300
C="1048576,1024,s-logistic,-1;2097152,512,s-hinge,-1;4194304,256,s-logistic,-1"
: > gpurun_out/sustained_wait.jsonl
for ns in 0 200 1000 5000; do
  NB200_WAIT_NS=$ns timeout 200 python tools/bench_sustained.py --seconds 2.5 --configs "$C" > gpurun_out/sw.jsonl 2>gpurun_out/sw.err
  python - $ns <<'PY'
import json,sys
for l in open('gpurun_out/sw.jsonl'):
    d=json.loads(l); d['wait_ns']=int(sys.argv[1]); open('gpurun_out/sustained_wait.jsonl','a').write(json.dumps(d)+'\n')
    print('wait_ns',sys.argv[1], d['features'], d['loss'], 'first', d['first_batch_us'], d['first_gbs'], 'steady', d['steady_us'], d['steady_gbs'], d['steady_sm_mhz'], d['steady_power_w'])
PY
done
