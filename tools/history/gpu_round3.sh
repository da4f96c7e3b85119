#!/bin/bash
# Third GPU visit: DMMA many-target kernels + solver e2e tests + fp64 ceilings.
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -o timeout_method=thread --maxfail=40 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peaks tools/fp64_peaks.cu && timeout 120 /tmp/fp64_peaks > gpurun_out/fp64_peaks.jsonl 2>&1; echo "peaks rc=$?"
cat gpurun_out/fp64_peaks.jsonl
timeout 600 python tools/bench_config3.py --rows 1048576 > gpurun_out/config3_1m.json 2>gpurun_out/config3_1m.err; echo "config3 1M rc=$?"
cat gpurun_out/config3_1m.json; tail -3 gpurun_out/config3_1m.err
timeout 900 python tools/bench_config3.py --rows 4194304 --no-cublas > gpurun_out/config3_4m.json 2>gpurun_out/config3_4m.err; echo "config3 4M rc=$?"
cat gpurun_out/config3_4m.json; tail -3 gpurun_out/config3_4m.err
timeout 600 python tools/bench_config3.py --rows 262144 --generic --no-cublas --steps 2 > gpurun_out/config3_generic.json 2>&1; echo "config3 generic rc=$?"
cat gpurun_out/config3_generic.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
timeout 300 python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/config3_small.json 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mt_ -s 2 -c 2 -f -o gpurun_out/prof_mt \
    python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas > gpurun_out/ncu_mt.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_mt.log
