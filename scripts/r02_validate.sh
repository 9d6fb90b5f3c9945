#!/usr/bin/env bash
# bring-up cases, the full GPU suite, smoke, then the C5 bench (value + e2e + eval) of the default build
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py pair > gpurun_out/bringup_pair.log 2>&1; echo "bringup rc=$?"; grep -E "worst|stuck" gpurun_out/bringup_pair.log
timeout 120 python tests/pair_timing.py 2>&1 | tail -n 1
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/suite.log 2>&1; echo "suite rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/suite.log | tail -n 30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 400 python bench.py --steps 100 --warmup 3 > gpurun_out/bench_C5.json 2> gpurun_out/bench_C5.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/bench_C5.json').read().strip().splitlines()[-1]); print('C5 ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value'] or 0), 'eval', round(d['eval_step']['ms_per_step'],4) if d.get('eval_step') and 'ms_per_step' in d['eval_step'] else d.get('eval_step'), 'roof', round(d['roofline']['frac'],4), [round(r['frac'],4) for r in d['roofline_other']], {k: round(v['ms_avg']*v['launches_per_step'],4) for k,v in d['kernels'].items()})" 2>&1 | tail -n 3
