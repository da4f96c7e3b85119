#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/bin/power_probe 3 > gpurun_out/power_probe.jsonl 2>gpurun_out/power_probe.err
cat gpurun_out/power_probe.jsonl; tail -3 gpurun_out/power_probe.err
