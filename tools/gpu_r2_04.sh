#!/bin/bash
# round 2, GPU call 4: two vs three stages in flight across the widths (single-target kernel)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out/r2_04_stages_sweep.jsonl
: > $OUT
CFG="1048576,1024,s-logistic,37;1048576,1024,mse,37;8388608,1024,s-logistic,37;2097152,512,s-hinge,33;2097152,512,s-logistic,33;4194304,256,s-logistic,34;8388608,128,s-logistic,35;16777216,64,s-logistic,36;524288,2048,s-logistic,39;1000000,1000,s-logistic,37;1048576,1024,s-logistic,27"
for ST in 2 3; do
  echo "{\"stage_bytes\": 65536, \"stages\": $ST}" >> $OUT
  NB200_STAGES=$ST timeout 600 python tools/bench_sustained.py --seconds 2.0 --configs "$CFG" >> $OUT 2>> gpurun_out/r2_04.err
done
echo "{\"stage_bytes\": 98304, \"stages\": 2}" >> $OUT
NB200_STAGE_BYTES=98304 NB200_STAGES=2 timeout 600 python tools/bench_sustained.py --seconds 2.0 --configs "$CFG" >> $OUT 2>> gpurun_out/r2_04.err
echo "{\"stage_bytes\": 32768, \"stages\": 4}" >> $OUT
NB200_STAGE_BYTES=32768 NB200_STAGES=4 timeout 600 python tools/bench_sustained.py --seconds 2.0 --configs "2097152,512,s-logistic,33;4194304,256,s-logistic,34;8388608,128,s-logistic,35" >> $OUT 2>> gpurun_out/r2_04.err
python - <<'PY'
import json
geo=None
for line in open('gpurun_out/r2_04_stages_sweep.jsonl'):
    d=json.loads(line)
    if 'stage_bytes' in d and 'kernel' not in d: geo=d; continue
    k=d['kernel']
    print(geo['stage_bytes'],geo['stages'],d['rows'],d['features'],d['loss'],k.split()[1],k[k.index('rows_per_stage'):], 'burst',d['burst_us_median'],d['burst_gbs'],'steady',d['steady_us'],d['steady_gbs'],d['steady_sm_mhz'])
PY
tail -3 gpurun_out/r2_04.err
