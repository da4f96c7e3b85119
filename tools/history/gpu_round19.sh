#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/probe_power.py > gpurun_out/probe_power.jsonl 2>gpurun_out/probe_power.err
cut -c1-420 gpurun_out/probe_power.jsonl; tail -3 gpurun_out/probe_power.err
nvidia-smi -q -d POWER | head -40
