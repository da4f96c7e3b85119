#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q --timeout 240 -o timeout_method=thread -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
{
  for loss in s-logistic s-hinge mse; do
    echo "# 2^20 x 1024 $loss"; timeout 300 python tools/sweep_t1.py --loss $loss --variants 4 --stage-bytes 65536 --stages 3
    echo "# 2^21 x 512 $loss: RW=2 (3) RW=1 (22) RW=4 (23)"; timeout 300 python tools/sweep_t1.py --rows 2097152 --features 512 --loss $loss --variants 3,22,23 --stage-bytes 65536 --stages 3
    echo "# 2^22 x 256 $loss: RW=4 (2) RW=1 (24) RW=8 (25)"; timeout 300 python tools/sweep_t1.py --rows 4194304 --features 256 --loss $loss --variants 2,24,25 --stage-bytes 65536 --stages 3
    echo "# 2^23 x 128 $loss: RW=8 (1) RW=1 (26)"; timeout 300 python tools/sweep_t1.py --rows 8388608 --features 128 --loss $loss --variants 1,26 --stage-bytes 65536 --stages 3
  done
  echo "# 2^24 x 64 logistic: RW=8 (0)"; timeout 300 python tools/sweep_t1.py --rows 16777216 --features 64 --variants 0 --stage-bytes 65536 --stages 3
} > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
for line in open('gpurun_out/sweep.log'):
    line=line.strip()
    if line.startswith('#'): print(line)
    elif line.startswith('{'):
        d=json.loads(line); print('   variant', d.get('variant'), d.get('plan','')[:48].split('grid')[0], 'ms_med', d.get('ms_med'), 'GB/s', d.get('gbs_med'), d.get('error',''))
PY
