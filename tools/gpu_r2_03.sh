#!/bin/bash
# round 2, GPU call 3: why is the cheap-loss pass slower at full clocks? ring geometry sweep (existing env knobs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out/r2_03_ring_sweep.jsonl
: > $OUT
CFG="1048576,1024,mse,27;1048576,1024,s-logistic,37;1048576,1024,s-logistic,27"
for SB in 16384 32768 65536; do
  for ST in 2 3 4 6 8; do
    echo "{\"stage_bytes\": $SB, \"stages\": $ST}" >> $OUT
    NB200_STAGE_BYTES=$SB NB200_STAGES=$ST timeout 300 python tools/bench_sustained.py --seconds 0.8 --configs "$CFG" >> $OUT 2>> gpurun_out/r2_03.err
  done
done
python - <<'PY'
import json
geo=None
for line in open('gpurun_out/r2_03_ring_sweep.jsonl'):
    d=json.loads(line)
    if 'stage_bytes' in d: geo=d; continue
    k=d['kernel']
    print(geo['stage_bytes'],geo['stages'],d['loss'],k.split()[1],k[k.index('rows_per_stage'):], 'burst',d['burst_us_median'],'steady',d['steady_us'],d['steady_sm_mhz'])
PY
timeout 600 python -m pytest tests/test_gpu_gboost.py -x -q -m gpu 2>&1 | tail -5
