#!/bin/bash
# e2e breakdown: where the host-buffer call spends its time (mapped-host stores vs D2H copy; own vs torch stream)
mkdir -p gpurun_out
: > gpurun_out/e2e_breakdown.jsonl
for copy in 0 1; do
  NB200_RESULT_COPY=$copy timeout 200 python tools/bench_e2e.py >> gpurun_out/e2e_breakdown.jsonl 2>gpurun_out/e2e.err
  NB200_RESULT_COPY=$copy timeout 200 python tools/bench_e2e.py --torch-stream >> gpurun_out/e2e_breakdown.jsonl 2>>gpurun_out/e2e.err
  NB200_RESULT_COPY=$copy timeout 200 python tools/bench_e2e.py --rows 10240 --features 1024 --calls 2000 >> gpurun_out/e2e_breakdown.jsonl 2>>gpurun_out/e2e.err
done
cut -c1-330 gpurun_out/e2e_breakdown.jsonl; tail -3 gpurun_out/e2e.err
