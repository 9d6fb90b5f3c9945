#!/usr/bin/env bash
# usage: r02_multi.sh <n_gpus> <config> [extra bench args]   (run under gpurun --gpus N)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; C=$2; shift 2
out=gpurun_out/bench_${C}_n${N}.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --config $C --steps 200 --warmup 3 --no-dense-probe --no-cpu "$@" > $out 2> gpurun_out/bench_${C}_n${N}.err; echo "rc=$?"
tail -n 3 gpurun_out/bench_${C}_n${N}.err | cut -c1-300
python -c "import json; d=json.loads(open('$out').read().strip().splitlines()[-1]); print('$C n=$N ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), 'e2e', d['e2e'] and round(d['e2e']['value'] or 0), 'eval', d.get('eval_step') and round(d['eval_step'].get('ms_per_step', 0),4), 'reshard', d.get('epoch_reshard') and round(d['epoch_reshard']['amortised_ms_per_step'],4), d['clocks'])" 2>&1 | tail -n 1
