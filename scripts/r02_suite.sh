#!/usr/bin/env bash
# the full GPU suite (all failures listed, no -x) + smoke
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/suite.log 2>&1; echo "suite rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/suite.log | tail -n 40
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
