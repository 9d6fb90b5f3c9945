#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
SCDEF_EXTRA_FLAGS=-DSCDEF_TC_TIMING bash scdef_b200/csrc/build.sh > /dev/null 2>&1
timeout 120 python tests/pair_timing.py 20000 100 dump 2>&1 | grep -E "kernel|rank 0" | cut -c1-600
