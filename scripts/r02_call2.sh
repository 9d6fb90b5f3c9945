#!/usr/bin/env bash
# Round 2, GPU call 2: the full GPU suite of the new default build, then the named BASELINE shapes on one GPU.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
t() { echo "=== $* ($(date +%T))" | tee -a $out/r02b_summary.log; }
t "1. full GPU suite (default build: pair mode 3 + k_sample_v2)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > $out/r02b_suite.log 2>&1; echo "rc=$?" | tee -a $out/r02b_summary.log
tail -n 30 $out/r02b_suite.log | tee -a $out/r02b_summary.log
t "2. smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee -a $out/r02b_summary.log
t "3. bench C2 / C3 (one GPU, S=100)"
for c in C2 C3; do
  timeout 300 python bench.py --config $c --steps 100 --warmup 3 --no-dense-probe > $out/r02b_bench_$c.json 2>> $out/r02b_summary.log
  python -c "import json; d=json.loads(open('gpurun_out/r02b_bench_$c.json').read().strip().splitlines()[-1]); print('$c ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value'] or 0), 'eval', d['eval_step'], {k: round(v['ms_avg']*v['launches_per_step'],4) for k,v in d['kernels'].items()})" 2>&1 | tee -a $out/r02b_summary.log
done
t "4. bench C5 default (one GPU)"
timeout 400 python bench.py --steps 100 --warmup 3 > $out/r02b_bench_C5.json 2>> $out/r02b_summary.log
python -c "import json; d=json.loads(open('gpurun_out/r02b_bench_C5.json').read().strip().splitlines()[-1]); print('C5 ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value'] or 0), 'eval', d['eval_step'], 'roof', d['roofline']['frac'], [r['frac'] for r in d['roofline_other']], {k: round(v['ms_avg']*v['launches_per_step'],4) for k,v in d['kernels'].items()})" 2>&1 | tee -a $out/r02b_summary.log
t "done"
