#!/bin/bash
# compute-sanitizer over the solver paths: memcheck on every case, racecheck + synccheck on the kernels that use
# named barriers, mbarrier, cp.async rings and the published-state counter (k_panel4, k_level_ws, k_spine, k_lin_gp).
# usage: scripts/sanitize.sh <tag>   -> gpurun_out/<tag>_{memcheck,racecheck,synccheck}.log
TAG=${1:-san}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 420 $CS --tool memcheck --leak-check no --error-exitcode 9 python scripts/sanitize_cases.py all 2 > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 420 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/sanitize_cases.py pose3_wide 1 > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck pose3_wide rc=$?"
timeout 420 $CS --tool racecheck --racecheck-report all --error-exitcode 9 python scripts/sanitize_cases.py sharded 1 > gpurun_out/${TAG}_racecheck_sharded.log 2>&1; echo "racecheck sharded rc=$?"
timeout 420 $CS --tool synccheck --error-exitcode 9 python scripts/sanitize_cases.py pose3_wide_loops 1 > gpurun_out/${TAG}_synccheck.log 2>&1; echo "synccheck rc=$?"
for f in gpurun_out/${TAG}_*.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_cases ok|Error|hazard" $f | sort | uniq -c | head -8; done
