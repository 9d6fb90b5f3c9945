#!/usr/bin/env bash
# round-2 closing measurements on one GPU: the named shapes, then the launch list and the --set full capture of the fused kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, args...
  local n=$1; shift
  timeout 600 python bench.py --steps 200 --warmup 3 --no-dense-probe "$@" > gpurun_out/final_$n.json 2> gpurun_out/final_$n.err; echo "$n rc=$?"
  python -c "import json; d=json.loads(open('gpurun_out/final_$n.json').read().strip().splitlines()[-1]); print('  $n ms/step', round(d['ms_per_step'],4), 'value', round(d['value']), 'e2e', d.get('e2e') and round(d['e2e']['value'] or 0), 'eval', d.get('eval_step') and round(d['eval_step'].get('ms_per_step', 0),4), 'frac', round(d['roofline']['frac'],4), 'cpu', d.get('cpu_baseline') and d['cpu_baseline'].get('value'))" 2>&1 | tail -n 1
}
run C5
run C5_S10 --samples 10 --no-cpu
run C5_S1 --samples 1 --no-cpu
run C2 --config C2 --no-cpu
run C3 --config C3 --no-cpu
run C4 --config C4 --no-cpu
bash scripts/r02_ncu.sh
python scripts/ncu_traffic.py gpurun_out/r02_lik_flat.ncu-rep C5 100 gpurun_out/r02_ncu_full_lik_flat.txt 2>&1 | tail -n 3
python profiles/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches.txt; head -32 gpurun_out/r02_launches.txt
