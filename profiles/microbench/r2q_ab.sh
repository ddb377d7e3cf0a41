#!/bin/bash
# Same-box A/B of library variants through bench.py: r2q_ab.sh <tag> <lib1> <lib2> ... (paths relative to the repo root)
cd "$(dirname "$0")/../.."
TAG=$1; shift
for rep in 1 2; do
for L in "$@"; do
  n=$(basename $L .so)
  SRL_B200_LIB=$PWD/$L timeout 600 python bench.py --steps 1000 --warmup 10 --e2e-steps 5 --no-cpu-baseline --no-extras > gpurun_out/r2q_${TAG}_${n}_$rep.json 2> gpurun_out/r2q_${TAG}_${n}_$rep.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2q_${TAG}_${n}_$rep.json').read().strip().splitlines()[-1])
k=d['kernels']
print('%-18s rep $rep: step %.2f us (warm %.2f) frac %.3f | K4 %.2f us K2 %.2f us | trainer-order %.1f us | e2e %.3g (resident %.3g) | parity %s' % (
  '$n', d['ms_per_step']*1e3, d['step']['ms_per_step_l2_warm']*1e3, d['step']['frac_of_peak'], k['ppo_loss_kernel']['ms_per_launch']*1e3,
  k['gae_scan_kernel']['ms_per_launch']*1e3, d['step_trainer_order']['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], d['parity_check']['ok']))
PY
done; done
