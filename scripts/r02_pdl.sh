#!/usr/bin/env bash
# programmatic dependent launch on / off: the GPU suite with it on, then step times at S = 100, 10, 1
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 3
for S in 100 10 1; do
  for off in 0 1; do
    if [ $off = 1 ]; then export SCDEF_NO_PDL=1; else unset SCDEF_NO_PDL; fi
    python bench.py --steps 300 --warmup 3 --samples $S --no-cpu --no-dense-probe --no-eval > gpurun_out/pdl_S${S}_off${off}.json 2>/dev/null
    python -c "import json; d=json.loads(open('gpurun_out/pdl_S${S}_off${off}.json').read().strip().splitlines()[-1]); print('S=$S no_pdl=$off  ms/step', round(d['ms_per_step'],4), ' e2e ms/step', round(d['e2e']['ms_per_step'],4))"
  done
done
