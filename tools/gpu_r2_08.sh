#!/bin/bash
# round 2, GPU call 8 (one GPU): DMMA issue-rate probe; the suite after the OSGA / fit / RQB changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 tools/bin/dmma_issue > gpurun_out/r2_08_dmma_issue.jsonl 2>&1; cat gpurun_out/r2_08_dmma_issue.jsonl
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_08_all_tests.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/r2_08_all_tests.log
