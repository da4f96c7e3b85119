#!/bin/bash
# ncu --set full captures of the tensor-pipe kernels (config 3 pair, one-pass fd) and the CUDA-core one-pass kernel
mkdir -p gpurun_out
cap() { # name regex command...
  local name=$1 regex=$2; shift 2
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 2 -f -o /tmp/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  ncu -i /tmp/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_${name}_raw.csv 2>/dev/null; echo "$name rc=$? $(wc -c < gpurun_out/ncu_${name}_raw.csv) bytes"
}
cap mt "mt_forward|mt_backward" python tools/bench_config3.py --rows 262144 --steps 2 --no-cublas
cap fd "vgrad_fd" python tools/bench_config3.py --rows 342392 --features 784 --targets 10 --steps 2 --no-cublas
cap ft "vgrad_ft" python tools/bench_config3.py --rows 262144 --features 1024 --targets 4 --steps 2 --no-cublas
