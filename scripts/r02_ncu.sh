#!/usr/bin/env bash
# launch list (per-kernel durations) and one --set full capture of the fused likelihood kernels, C5 shapes.
# Numbers printed by these profiled runs are never bench values.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-preroll --no-e2e --no-cpu --no-dense-probe --no-eval --cells 200000"
timeout 300 $B > gpurun_out/ncu_plain.json 2> gpurun_out/ncu_plain.err; echo "plain rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lik_flat --launch-skip 6 -c 2 -o gpurun_out/r02_lik_flat -f $B > gpurun_out/ncu2.log 2>&1; echo "full capture rc=$?"
ls -la gpurun_out/r02_lik_flat.ncu-rep gpurun_out/r02_launches.csv
