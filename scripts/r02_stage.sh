#!/usr/bin/env bash
# host placement (e2e): staging workers / slots / depth at S = 10 and S = 100
cd "$(dirname "$0")/.."
for S in 10 100; do
  for cfg in "1 3 2" "2 5 3" "2 6 4"; do
    set -- $cfg
    SCDEF_STAGE_WORKERS=$1 SCDEF_STAGE_SLOTS=$2 SCDEF_STAGE_DEPTH=$3 python bench.py --steps 300 --warmup 3 --samples $S --no-cpu --no-dense-probe --no-eval 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('S=$S workers/slots/depth=$cfg  value ms', round(d['ms_per_step'],4), ' e2e ms', round(d['e2e']['ms_per_step'],4), round(d['e2e']['value']))"
  done
done
