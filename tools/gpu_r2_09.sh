#!/bin/bash
# round 2, GPU call 9 (one GPU): is the many-target forward limited by its HBM feed? (tile rows forced L2-hot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for HOT in 0 1; do
  for ST in 0 2; do
    echo "hot=$HOT stages=$ST"
    NB200_MT_PROBE_HOT=$HOT NB200_MT_STAGES=$ST timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['plan'][:60], 'value ms', d['value']['ms_med'], d['value']['tflops'], 'vgrad ms', d['vgrad']['ms_med'])"
  done
done
timeout 900 python -m pytest tests/test_gpu_batch_trials.py -x -q -m gpu 2>&1 | tail -15
