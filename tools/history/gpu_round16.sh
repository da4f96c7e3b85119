#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py tests/test_gpu_parity.py -m gpu -q --timeout 240 -o timeout_method=thread -x -k "batch or many_target or multi_target or extreme or invalid" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log | cut -c1-300
for st in 0 3; do
  NB200_MT_STAGES=$st timeout 900 python tools/bench_config3.py --rows 4194304 --no-cublas > gpurun_out/config3_4m_$st.json 2>gpurun_out/config3_4m.err; echo "config3 4M stages=$st rc=$?"
  cat gpurun_out/config3_4m_$st.json | cut -c1-600; tail -3 gpurun_out/config3_4m.err
done
