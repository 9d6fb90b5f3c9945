#!/usr/bin/env bash
# one-GPU measurements of the final build: named shapes (C2, C3, C5 at S = 100 / 10 / 1), launch list, full ncu capture
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, args...
  local name=$1; shift
  timeout 500 python bench.py "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "$name rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/bench_$name.json').read().strip().splitlines()[-1]); print('$name ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value'] or 0), 'eval', round(d['eval_step']['ms_per_step'],4) if d.get('eval_step') and 'ms_per_step' in d['eval_step'] else None, 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']), 'roof', [round(r['frac'],4) for r in [d['roofline']]+d['roofline_other']], d['clocks'])" 2>&1 | tail -n 1
}
run C5 --steps 200 --warmup 3
run C5_S10 --samples 10 --steps 200 --warmup 3 --no-cpu --no-dense-probe
run C5_S1 --samples 1 --steps 200 --warmup 3 --no-cpu --no-dense-probe
run C2 --config C2 --steps 200 --warmup 3 --no-dense-probe
run C3 --config C3 --steps 200 --warmup 3 --no-dense-probe
B="python bench.py --steps 3 --warmup 3 --no-preroll --no-e2e --no-cpu --no-dense-probe --no-eval --cells 200000"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/r02_launches_final.csv $B > gpurun_out/ncu1.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lik_flat --launch-skip 6 -c 2 -o gpurun_out/r02_lik_flat_final -f $B > gpurun_out/ncu2.log 2>&1; echo "full capture rc=$?"
