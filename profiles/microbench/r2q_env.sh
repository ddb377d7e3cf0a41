#!/bin/bash
# Same-box A/B of environment settings through bench.py: r2q_env.sh <tag> "<ENV=.. ENV=..>" "<...>" ...   ("-" = no setting)
cd "$(dirname "$0")/../.."
TAG=$1; shift
for rep in 1 2; do
i=0
for E in "$@"; do
  i=$((i+1)); [ "$E" = "-" ] && E=""
  env $E timeout 600 python bench.py --steps 1000 --warmup 10 --e2e-steps 5 --no-cpu-baseline --no-extras > gpurun_out/r2q_${TAG}_${i}_$rep.json 2> gpurun_out/r2q_${TAG}_${i}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2q_${TAG}_${i}_$rep.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print('%-28s rep $rep: step %.2f us (warm %.2f) frac %.3f | K4 %.2f us K2 %.2f us | trainer-order %.1f us | e2e %.3g (resident %.3g) | parity %s launches %s' % (
      '[$E]', d['ms_per_step']*1e3, d['step']['ms_per_step_l2_warm']*1e3, d['step']['frac_of_peak'], k['ppo_loss_kernel']['ms_per_launch']*1e3,
      k['gae_scan_kernel']['ms_per_launch']*1e3, d['step_trainer_order']['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], d['parity_check']['ok'], d['gpu_launches_per_step']))
except Exception as e:
    print('[$E] FAILED', e); print(open('gpurun_out/r2q_${TAG}_${i}_$rep.err').read()[-600:])
PY
done; done
