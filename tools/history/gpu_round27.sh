#!/bin/bash
# config 1 latency with the L2-resident policy; ncu launch list + full capture of the hot kernel (variant 27); config 3
mkdir -p gpurun_out
timeout 200 python tools/bench_config1.py > gpurun_out/config1.jsonl 2>gpurun_out/config1.err; echo "config1 rc=$?"; cut -c1-400 gpurun_out/config1.jsonl; tail -2 gpurun_out/config1.err
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -o timeout_method=thread -k "t1 or synth or builtin" > gpurun_out/pytest_small.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_small.log | cut -c1-200
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:vgrad_t1 -s 4 -c 2 -f -o /tmp/prof_t1 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 && \
ncu -i /tmp/prof_t1.ncu-rep --page raw --csv > gpurun_out/ncu_t1_raw.csv 2>/dev/null; \
ncu -i /tmp/prof_t1.ncu-rep --page source --csv > gpurun_out/ncu_t1_source.csv 2>/dev/null
echo "ncu rc=$?"; ls -la gpurun_out/launches.csv gpurun_out/ncu_t1_raw.csv
timeout 300 python tools/bench_config3.py --rows 1048576 --no-cublas > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 rc=$?"; cut -c1-700 gpurun_out/config3_1m.json
