#!/bin/bash
# round 2, GPU call 17: many-target backward with a predicate-free fast path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu 2>&1 | tail -2
for PROBE in 0 2; do
  NB200_MT_PROBE_BWD=$PROBE timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('probe $PROBE', 'value ms', d['value']['ms_med'], 'vgrad ms', d['vgrad']['ms_med'], d['vgrad']['tflops'], 'backward ms', round(d['vgrad']['ms_med'] - d['value']['ms_med'], 3))"
done
