#!/bin/bash
# round 2, GPU call 13: many-target forward, epilogue of four rows at a time on the consumer warps (tensor-map fed / cp.async fed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('tma   ', d['plan'][:60], 'value ms', d['value']['ms_med'], d['value']['tflops'], 'vgrad ms', d['vgrad']['ms_med'], d['vgrad']['tflops'])"
NB200_MT_LEGACY_FORWARD=1 timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 --steps 4 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('legacy', d['plan'][:60], 'value ms', d['value']['ms_med'], d['value']['tflops'], 'vgrad ms', d['vgrad']['ms_med'], d['vgrad']['tflops'])"
timeout 300 python tools/bench_batch.py 2>&1 | tail -12
