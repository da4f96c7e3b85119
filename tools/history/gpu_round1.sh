#!/bin/bash
# First GPU visit: smoke, parity suite, variant sweep, bench, ncu launch list (only if the plain bench run passed).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -5 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_t1.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
