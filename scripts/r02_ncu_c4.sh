#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --config C4 --steps 3 --warmup 3 --no-preroll --no-e2e --no-cpu --no-dense-probe --no-eval --cells 100000"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches_c4.csv $B > gpurun_out/ncu_c4.log 2>&1; echo "launch list rc=$?"
python profiles/summarize_launches.py gpurun_out/r02_launches_c4.csv | head -30
