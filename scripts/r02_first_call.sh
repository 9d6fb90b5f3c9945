#!/usr/bin/env bash
# Everything that was written after round 1's GPU minutes were spent, in ONE gpurun call (about 13 minutes of box time):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/r02_first_call.sh'
# Results land in gpurun_out/r02_*.log.  Nothing here can hang the device: the pair kernels and the probes bound their waits.
set -u
cd "$(dirname "$0")/.."
out=gpurun_out
mkdir -p $out
t() { echo "=== $* ($(date +%T))" | tee -a $out/r02_summary.log; }

t "1. pending GPU tests (assign_confident kernel, device init, pair mode 3)"
timeout 300 python -m pytest tests -q -m gpu_pending -p no:cacheprovider > $out/r02_pending.log 2>&1; echo "rc=$?" | tee -a $out/r02_summary.log
tail -n 5 $out/r02_pending.log | tee -a $out/r02_summary.log

t "2. the full GPU suite with the persistent CTA-pair LOCAL kernel as the default path"
SCDEF_TC_PAIR=2 timeout 420 python -m pytest tests -q -m gpu -p no:cacheprovider > $out/r02_suite_pair2.log 2>&1; echo "rc=$?" | tee -a $out/r02_summary.log
tail -n 5 $out/r02_suite_pair2.log | tee -a $out/r02_summary.log

t "3. per-launch time of the likelihood kernels: single CTA, pair persistent, + whole-pair-tile epilogue, + per-k-step Q hand-off"
for mode in 0 2 3 4 5; do  # SCDEF_SAMPLE_V2=1 in front of a run times k_sample_v2 as well
  SCDEF_TC_PAIR=$mode timeout 60 python tests/pair_timing.py 2>&1 | tail -n 1 | tee -a $out/r02_summary.log
done

SCDEF_TC_L2PF=1 timeout 60 python tests/pair_timing.py 2>&1 | tail -n 1 | sed 's/^/[L2 prefetch of W images] /' | tee -a $out/r02_summary.log

SCDEF_TC_POLL=1 timeout 60 python tests/pair_timing.py 2>&1 | tail -n 1 | sed 's/^/[MMA thread polls operand readiness] /' | tee -a $out/r02_summary.log
SCDEF_TC_POLL=1 SCDEF_TC_L2PF=1 timeout 60 python tests/pair_timing.py 2>&1 | tail -n 1 | sed 's/^/[poll + L2 prefetch] /' | tee -a $out/r02_summary.log
SCDEF_TC_POLL=1 SCDEF_TC_L2PF=1 timeout 120 python tests/tc_bringup.py > $out/r02_bringup_poll.log 2>&1; grep "tc:" $out/r02_bringup_poll.log | tee -a $out/r02_summary.log

t "4. semantics probes for the GLOBAL pair kernel (one question per process)"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/mma_probe3 scripts/mma_probe3.cu 2>> $out/r02_summary.log
for q in 1 2 3; do timeout 30 scripts/mma_probe3 $q 2>&1 | tee -a $out/r02_summary.log; done

t "5. whole train step at C5: default build, then with the opt-in variants (pair persistent LOCAL kernel + k_sample_v2)"
timeout 240 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu --no-dense-probe > $out/r02_bench_default.json 2>> $out/r02_summary.log
SCDEF_TC_PAIR=2 SCDEF_SAMPLE_V2=1 timeout 240 python bench.py --steps 50 --warmup 3 --no-e2e --no-cpu --no-dense-probe > $out/r02_bench_variants.json 2>> $out/r02_summary.log
python - <<'PY' | tee -a $out/r02_summary.log
import json
for name in ("default", "variants"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{name}.json").read().strip().splitlines()[-1])
        k = d["kernels"]
        print(name, "ms/step", round(d["ms_per_step"], 4), {n: round(v["ms_avg"] * v["launches_per_step"], 4) for n, v in k.items()})
    except Exception as e:  # noqa: BLE001
        print(name, "no result:", e)
PY

t "6. secondary settings of SURVEY 8(d): S = 10 (paper benchmarks) and S = 1, default build"
for S in 10 1; do
  timeout 200 python bench.py --samples $S --steps 50 --warmup 3 --no-e2e --no-cpu --no-dense-probe > $out/r02_bench_S$S.json 2>> $out/r02_summary.log
  python -c "import json,sys; d=json.loads(open('gpurun_out/r02_bench_S$S.json').read().strip().splitlines()[-1]); print('S=$S ms/step', round(d['ms_per_step'],4), 'cells/s', round(d['value']))" 2>&1 | tee -a $out/r02_summary.log
done

t "done"
