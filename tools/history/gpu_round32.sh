#!/bin/bash
# bisect on one box: config 3 (1 Mi rows) with the libraries of three earlier commits and the current build
mkdir -p gpurun_out
for lib in tools/bin/libnb200_4c89a21.so tools/bin/libnb200_408b402.so tools/bin/libnb200_82c0247.so libnano_b200/libnb200.so; do
  NB200_LIBRARY=$PWD/$lib timeout 200 python tools/bench_config3.py --rows 1048576 --no-cublas --steps 6 > gpurun_out/c3_ab.json 2>gpurun_out/c3.err; echo "$lib rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c3_ab.json')); print(d['vgrad']['ms_med'], d['value']['ms_med'], d['plan'])"; tail -2 gpurun_out/c3.err
done
