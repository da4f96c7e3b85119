#!/bin/bash
# Second GPU visit: fused-tail kernel. Parity suite, sweeps, bench, then ncu (launch list + full capture of the hot kernel).
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -o timeout_method=thread --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
{
  echo "# 2^20 x 1024 logistic"; timeout 600 python tools/sweep_t1.py --variants 4,17,20 --stage-bytes 65536 --stages 2,3
  echo "# 2^20 x 1024 logistic value-only"; timeout 300 python tools/sweep_t1.py --variants 4 --stage-bytes 65536 --stages 3 --value-only
  echo "# 2^20 x 1024 mse"; timeout 300 python tools/sweep_t1.py --variants 4 --stage-bytes 65536 --stages 3 --loss mse
  echo "# 2^21 x 512 hinge"; timeout 600 python tools/sweep_t1.py --rows 2097152 --features 512 --loss s-hinge --variants 3,17 --stage-bytes 32768,65536 --stages 3,6
  echo "# 2^19 x 2048 logistic"; timeout 600 python tools/sweep_t1.py --rows 524288 --features 2048 --variants 5 --stage-bytes 65536 --stages 3
  echo "# 2^22 x 256 logistic"; timeout 600 python tools/sweep_t1.py --rows 4194304 --features 256 --variants 2 --stage-bytes 32768,65536 --stages 3,6
} > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
cat gpurun_out/sweep.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"
cat gpurun_out/bench_reference.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:vgrad_t1 -s 4 -c 2 -f -o gpurun_out/prof_t1 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/ncu_full.log
