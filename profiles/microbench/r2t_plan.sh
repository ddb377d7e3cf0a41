#!/bin/bash
# Same-box A/B: the timed step issued as one CUDA graph replay or as the recorded plan of stream launches (r2t)
cd "$(dirname "$0")/../.."
TAG=${1:-r2t}; STEPS=${2:-600}
for rep in 1 2; do
for L in graph plan; do
  timeout 300 python bench.py --launch $L --steps $STEPS --warmup 10 --e2e-steps 5 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_${L}_$rep.json 2> gpurun_out/${TAG}_${L}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_${L}_$rep.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print('%-6s rep $rep: step %.2f us (warm %.2f; other launch %s) frac %.3f | K4 %.2f us K2 %.2f us | parity %s' % (
      '$L', d['ms_per_step']*1e3, d['step']['ms_per_step_l2_warm']*1e3, {k_: round(v*1e3, 2) for k_, v in d['step']['ms_per_step_other_launch'].items()},
      d['step']['frac_of_peak'], k['ppo_loss_kernel']['ms_per_launch']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, d['parity_check']['ok']))
except Exception as ex:
    print('$L rep $rep: FAILED', ex)
PY
done; done
