#!/usr/bin/env bash
# bring-up of the unified pair kernel (lik_flat.cuh): parity cases, per-launch timing, then the per-role cycle breakdown
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bring-up cases (tc vs simt vs oracle)"
timeout 300 python tests/tc_bringup.py > gpurun_out/bringup.log 2>&1; echo "rc=$?"
grep -E "case|tc:|Error|error|stuck" gpurun_out/bringup.log | cut -c1-260
echo "== pair cases"
timeout 300 python tests/tc_bringup.py pair > gpurun_out/bringup_pair.log 2>&1; echo "rc=$?"
grep -E "pair|tc:|Error|error|stuck" gpurun_out/bringup_pair.log | cut -c1-260
echo "== timing"
timeout 120 python tests/pair_timing.py 2>&1 | tail -n 1
if [ "${1:-}" = "dump" ]; then
  echo "== cycle breakdown (-DSCDEF_TC_TIMING build)"
  SCDEF_EXTRA_FLAGS=-DSCDEF_TC_TIMING bash scdef_b200/csrc/build.sh > /dev/null 2>&1
  timeout 120 python tests/pair_timing.py 20000 100 dump 2>&1 | tail -n 20
fi
