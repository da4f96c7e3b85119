#!/bin/bash
# host-driven RQB (proximal bundle) through the GPU vgrad: the new parity test, then config 5 in full (2^24 x 512)
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_solvers.py -m gpu -q -x -k "rqb" --timeout 80 -o timeout_method=thread > gpurun_out/pytest_rqb.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_rqb.log
timeout 100 python tools/solve_config.py --config 5 --nonsmooth-solvers rqb --max-evals 800 > gpurun_out/solve_config5_rqb.jsonl 2> gpurun_out/solve_config5_rqb.err
echo "solve rc=$?"
tail -3 gpurun_out/pytest_rqb.log
cat gpurun_out/solve_config5_rqb.jsonl | cut -c1-300
