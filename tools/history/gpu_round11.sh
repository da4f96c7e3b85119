#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q --timeout 240 -o timeout_method=thread -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
{
  echo "# 2^20 x 1024 logistic: RW=1 (4) vs RW=2 (21)"; timeout 300 python tools/sweep_t1.py --variants 4,21 --stage-bytes 65536 --stages 3
  echo "# 2^21 x 512 hinge: RW=2 (3) vs RW=1 (22) vs RW=4 (23)"; timeout 300 python tools/sweep_t1.py --rows 2097152 --features 512 --loss s-hinge --variants 3,22,23 --stage-bytes 65536 --stages 3
  echo "# 2^21 x 512 logistic"; timeout 300 python tools/sweep_t1.py --rows 2097152 --features 512 --variants 3,22,23 --stage-bytes 65536 --stages 3
  echo "# 2^22 x 256 logistic: RW=4 (2) vs RW=1 (24) vs RW=8 (25)"; timeout 300 python tools/sweep_t1.py --rows 4194304 --features 256 --variants 2,24,25 --stage-bytes 65536 --stages 3
  echo "# 2^23 x 128 logistic: RW=8 (1) vs RW=1 (26)"; timeout 300 python tools/sweep_t1.py --rows 8388608 --features 128 --variants 1,26 --stage-bytes 65536 --stages 3
  echo "# 2^23 x 100 mse (config-1 width)"; timeout 300 python tools/sweep_t1.py --rows 8388608 --features 100 --loss mse --variants 1,26 --stage-bytes 65536 --stages 3
  echo "# 2^24 x 64 logistic: RW=8 (0)"; timeout 300 python tools/sweep_t1.py --rows 16777216 --features 64 --variants 0 --stage-bytes 65536 --stages 3
  echo "# 2^20 x 1023 logistic (odd D, V=1)"; timeout 300 python tools/sweep_t1.py --rows 1048576 --features 1023 --variants 13 --stage-bytes 65536 --stages 3
} > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cut -c1-260 gpurun_out/sweep.log
timeout 300 python tools/bench_config1.py > gpurun_out/config1.jsonl 2>&1; cut -c1-300 gpurun_out/config1.jsonl
