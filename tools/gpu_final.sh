#!/bin/bash
# Round-end evidence: full GPU suite, both bench arms, ncu launch list + full capture of the hot kernel (CSV exports).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q --timeout 120 -o timeout_method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
cut -c1-500 gpurun_out/bench_reference.log
timeout 600 python bench.py --steps 50 --warmup 5 --sustained-seconds 3 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vgrad_t1 -s 4 -c 2 -f -o /tmp/prof_t1 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1 && \
ncu -i /tmp/prof_t1.ncu-rep --page raw --csv > gpurun_out/ncu_t1_raw.csv 2>/dev/null; \
ncu -i /tmp/prof_t1.ncu-rep --page source --csv > gpurun_out/ncu_t1_source.csv 2>/dev/null
echo "ncu rc=$?"; ls -la /tmp/prof_t1.ncu-rep gpurun_out/*.csv
