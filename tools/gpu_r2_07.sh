#!/bin/bash
# round 2, GPU call 5: the tensor-map forward of config 3 against the cp.async one; full GPU suite; bench; multi-target
# kernels with fewer stages in flight
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -x -q -m gpu -k "many or softmax or mt or multi or target or batch" > gpurun_out/r2_07_mt_tests.log 2>&1
echo "mt tests rc=$?"; tail -5 gpurun_out/r2_07_mt_tests.log
timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 > gpurun_out/r2_07_cfg3_tma_1m.json 2> gpurun_out/r2_07_cfg3.err
echo "cfg3 tma rc=$?"; cat gpurun_out/r2_07_cfg3_tma_1m.json; tail -3 gpurun_out/r2_07_cfg3.err
NB200_MT_LEGACY_FORWARD=1 timeout 300 python tools/bench_config3.py --no-cublas --rows 1048576 > gpurun_out/r2_07_cfg3_legacy_1m.json 2>> gpurun_out/r2_07_cfg3.err
cat gpurun_out/r2_07_cfg3_legacy_1m.json
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_07_all_tests.log 2>&1
echo "all tests rc=$?"; tail -5 gpurun_out/r2_07_all_tests.log
timeout 600 python bench.py > gpurun_out/r2_07_bench_n1.json 2> gpurun_out/r2_07_bench_n1.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_07_bench_n1.err
