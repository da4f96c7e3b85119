#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/probe_gap.jsonl
timeout 200 python tools/probe_gap.py >> gpurun_out/probe_gap.jsonl 2>gpurun_out/probe.err
timeout 200 python tools/probe_gap.py --torch-stream >> gpurun_out/probe_gap.jsonl 2>>gpurun_out/probe.err
NB200_PLAIN_LAUNCH=1 timeout 200 python tools/probe_gap.py >> gpurun_out/probe_gap.jsonl 2>>gpurun_out/probe.err
timeout 200 python tools/probe_gap.py --loss mse >> gpurun_out/probe_gap.jsonl 2>>gpurun_out/probe.err
cut -c1-200 gpurun_out/probe_gap.jsonl; tail -3 gpurun_out/probe.err
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,pstate --format=csv
