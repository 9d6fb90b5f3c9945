#!/usr/bin/env bash
# A/B of compile-time variants of the fused kernel: per-launch timing and the step time of the default bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
i=0
for f in "$@"; do
  i=$((i+1))
  SCDEF_EXTRA_FLAGS="$f" bash scdef_b200/csrc/build.sh > /dev/null 2>&1
  echo "flags [$f]: $(timeout 120 python tests/pair_timing.py 2>&1 | tail -n 1)"
  python bench.py --steps 200 --warmup 3 --no-cpu --no-dense-probe --no-eval --no-e2e > gpurun_out/ab_$i.json 2>/dev/null
  python -c "import json; d=json.loads(open('gpurun_out/ab_$i.json').read().strip().splitlines()[-1]); k=d['kernels']; print('   step ms', round(d['ms_per_step'],4), {n: round(k[n]['ms_avg'],4) for n in ('lik','lik_main','lik_main_global','hier')})"
done
