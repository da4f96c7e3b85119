#!/bin/bash
# One GPU: parity of the lane-loss kernels (vgrad_fl.cuh), then the A/B sweep against the kernels they replace.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [ -z "$FL_SKIP_TESTS" ]; then
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lane_loss or handful or feature_split" > gpurun_out/fl_tests.log 2>&1
echo "fl tests rc=$?"; tail -5 gpurun_out/fl_tests.log
fi
SHAPES="784,10,s-classnll;512,16,s-classnll;640,12,s-classnll;384,10,s-classnll;832,13,mse;1024,8,s-classnll;768,8,s-classnll;512,8,s-classnll;1000,5,s-classnll;512,6,m-logistic;320,4,s-classnll"
: > gpurun_out/fl_sweep.jsonl
IFS=';' read -ra MODES <<< "${FL_MODES:-0 0 1;1 0 1;2 0 2}"   # "NB200_FL NB200_FL_TPI NB200_FL_PIPE;..."
for mode in "${MODES[@]}"; do
  IFS=' ' read -r a b c <<< "$mode"; set -- $a $b $c
  echo "{\"NB200_FL\": $1, \"NB200_FL_TPI\": $2, \"NB200_FL_PIPE\": $3}" >> gpurun_out/fl_sweep.jsonl
  NB200_FL=$1 NB200_FL_TPI=$2 NB200_FL_PIPE=$3 timeout 120 python tools/sweep_multi.py --shapes "$SHAPES" --gib 2 --seconds 0.5 >> gpurun_out/fl_sweep.jsonl 2> gpurun_out/fl_sweep.err
  echo "sweep $mode rc=$?"
done
