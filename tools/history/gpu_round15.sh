#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/config3_small.json 2>&1 && \
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:mt_forward -s 1 -c 1 -f -o /tmp/prof_fwd \
    python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/ncu_fwd.log 2>&1 && \
ncu -i /tmp/prof_fwd.ncu-rep --page source --csv > gpurun_out/ncu_fwd_source.csv 2>/dev/null; \
ncu -i /tmp/prof_fwd.ncu-rep --page raw --csv > gpurun_out/ncu_fwd_raw.csv 2>/dev/null
echo "ncu rc=$?"; ls -la /tmp/prof_fwd.ncu-rep gpurun_out/ncu_fwd_source.csv
