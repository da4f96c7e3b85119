#!/bin/bash
# 8-GPU visit: NCCL at world 8 (check + bench at 8 and 4), kept short.
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded.py > gpurun_out/check_sharded_8gpu.json 2>gpurun_out/check_sharded_8gpu.err; echo "check_sharded(8) rc=$?"
tail -1 gpurun_out/check_sharded_8gpu.json | cut -c1-1200; tail -3 gpurun_out/check_sharded_8gpu.err | cut -c1-300
for n in 8 4 2 1; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${n}gpu.log 2>gpurun_out/bench_${n}gpu.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_${n}gpu.log 2>gpurun_out/bench_${n}gpu.err
  fi
  echo "bench($n) rc=$?"; tail -1 gpurun_out/bench_${n}gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'frac', d['roofline']['frac'], d['clocks'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/solve_config.py --config 4 > gpurun_out/solve_config4_8gpu.jsonl 2>gpurun_out/solve_config4_8gpu.err; echo "solve4(8 gpus, 64Mi rows) rc=$?"
cat gpurun_out/solve_config4_8gpu.jsonl | cut -c1-700; tail -3 gpurun_out/solve_config4_8gpu.err | cut -c1-300
