#!/bin/bash
# A/B of programmatic dependent launch (SRL_PDL) and the deep TMA ring of the warp-specialised scan (SRL_GAE_WS_DEEP)
# on the default bench workload; writes one bench line per combination to gpurun_out/ab_<tag>.json.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CFG=${1:-cfg2_atari_large}
for pdl in 0 1; do for deep in 0 1; do
  SRL_PDL=$pdl SRL_GAE_WS_DEEP=$deep python bench.py --config $CFG --steps 1000 --warmup 20 --e2e-steps 20 --no-cpu-baseline \
    > gpurun_out/ab_${CFG}_pdl${pdl}_deep${deep}.json 2> gpurun_out/ab_${CFG}_pdl${pdl}_deep${deep}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_${CFG}_pdl${pdl}_deep${deep}.json").read().strip().splitlines()[-1])
    k = d["kernels"]
    print("pdl=$pdl deep=$deep  step %.2f us (warm %.2f)  K2 %.2f us  K4 %.2f us  e2e %.0f us" % (
        d["ms_per_step"] * 1e3, d["step"]["ms_per_step_l2_warm"] * 1e3, k["gae_scan_kernel"]["ms_per_launch"] * 1e3,
        k["ppo_loss_kernel"]["ms_per_launch"] * 1e3, d["e2e"]["ms_per_step"] * 1e3))
except Exception as e:
    print("pdl=$pdl deep=$deep FAILED", e)
PY
done; done
