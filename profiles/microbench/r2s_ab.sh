#!/bin/bash
# Same-box A/B of the minibatch statistics prologue of the loss kernel (r2s):
#   gather : SRL_MB_PART=0 -- the problem's first CTA gathers its minibatch's 512 lane items through the permutation, publishes
#   each   : every loss CTA adds the scan's 128 per-CTA shares of its minibatch itself (product build)
#   share  : -DSRL_PAIR_PART_EACH=0 -- the first CTA adds the shares and publishes, the others poll
# usage: r2s_ab.sh <tag> [steps]
cd "$(dirname "$0")/../.."
TAG=${1:-r2s}; STEPS=${2:-600}
run() {  # name rep env...
  local n=$1 rep=$2; shift 2
  env "$@" timeout 300 python bench.py --steps $STEPS --warmup 10 --e2e-steps 5 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_${n}_$rep.json 2> gpurun_out/${TAG}_${n}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_${n}_$rep.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print('%-8s rep $rep: step %.2f us (warm %.2f) frac %.3f | K4 %.2f us K2 %.2f us | trainer-order %.1f us | e2e %.3g (resident %.3g) | parity %s' % (
      '$n', d['ms_per_step']*1e3, d['step']['ms_per_step_l2_warm']*1e3, d['step']['frac_of_peak'], k['ppo_loss_kernel']['ms_per_launch']*1e3,
      k['gae_scan_kernel']['ms_per_launch']*1e3, d['step_trainer_order']['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], d['parity_check']['ok']))
except Exception as ex:
    print('$n rep $rep: FAILED', ex)
PY
}
for rep in 1 2; do
  run gather $rep SRL_MB_PART=0
  run each $rep SRL_MB_PART=1
  run share $rep SRL_MB_PART=1 SRL_B200_LIB=$PWD/srl_b200/libsrl_b200_share.so
done
