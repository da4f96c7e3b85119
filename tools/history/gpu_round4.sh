#!/bin/bash
# Fourth GPU visit (1 GPU): DMMA forward with cp.async producer; config 3 numbers; small ncu csv.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 240 -o timeout_method=thread -x -k "many_target or multi_target" > gpurun_out/pytest_mt.log 2>&1; echo "pytest mt rc=$?"
tail -5 gpurun_out/pytest_mt.log | cut -c1-300
timeout 600 python tools/bench_config3.py --rows 1048576 --no-cublas > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 1M rc=$?"
cat gpurun_out/config3_1m.json; tail -3 gpurun_out/config3_1m.err
timeout 900 python tools/bench_config3.py --rows 4194304 --no-cublas > gpurun_out/config3_4m.json 2>gpurun_out/config3_4m.err; echo "config3 4M rc=$?"
cat gpurun_out/config3_4m.json; tail -3 gpurun_out/config3_4m.err
timeout 300 python tools/check_sharded.py > gpurun_out/check_sharded_1gpu.json 2>&1; echo "check_sharded(1) rc=$?"
tail -2 gpurun_out/check_sharded_1gpu.json | cut -c1-600
timeout 300 python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/config3_small.json 2>&1 && \
ncu --set full --clock-control none -k regex:mt_ -s 2 -c 2 -f -o /tmp/prof_mt \
    python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/ncu_mt.log 2>&1 && \
ncu -i /tmp/prof_mt.ncu-rep --page raw --csv > gpurun_out/ncu_mt_raw.csv 2>/dev/null
echo "ncu rc=$?"; ls -la /tmp/prof_mt.ncu-rep gpurun_out/ | head
