#!/bin/bash
# round 2, GPU call 16: what limits the many-target backward? (L2-hot feed; consumers free-running)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for PROBE in 0 1 2; do
  NB200_MT_PROBE_BWD=$PROBE timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('probe $PROBE', 'value ms', d['value']['ms_med'], 'vgrad ms', d['vgrad']['ms_med'], 'backward ms', round(d['vgrad']['ms_med'] - d['value']['ms_med'], 3))"
done
