#!/bin/bash
# round 2, GPU call 11: what does the stage protocol cost the many-target forward? (consumers free-running on stale tiles)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for HOT in 4 0; do
  echo "probe level $HOT"
  NB200_MT_PROBE_HOT=$HOT timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['plan'][:60], 'value ms', d['value']['ms_med'], d['value']['tflops'], 'vgrad ms', d['vgrad']['ms_med'])"
done
