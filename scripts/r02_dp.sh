#!/usr/bin/env bash
# data-parallel checks on N GPUs: the two-rank tests, then the bench with and without the overlapped all-reduce
# usage: r02_dp.sh <n_gpus> <config>
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$1; C=$2
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py -q -m gpu -k "lik_first or two_rank" 2>&1 | tail -n 3
for ov in 1 0; do
  out=gpurun_out/dp_${C}_n${N}_ov${ov}.json
  SCDEF_DP_OVERLAP=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --config $C --steps 300 --warmup 3 --no-dense-probe --no-cpu --no-eval --no-e2e > $out 2> gpurun_out/dp_${C}_n${N}_ov${ov}.err; echo "overlap=$ov rc=$?"
  python -c "import json; d=json.loads(open('$out').read().strip().splitlines()[-1]); print('  $C n=$N overlap=$ov ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']), d.get('step_ms'))" 2>&1 | tail -n 1
done
