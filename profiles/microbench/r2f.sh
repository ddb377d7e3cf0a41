#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {
  env SRL_B200_LIB=$PWD/srl_b200/$2 $3 timeout 300 python bench.py --steps 300 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('%-12s step %.1f us K2 %.1f K4 %.1f (warm %.1f) frac %.3f' % ('$1', d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['bytes_per_launch']/k['ppo_loss_kernel']['gbs_l2_warm']/1e3, d['step']['frac_of_peak']))"
}
for rep in 1 2; do
  run reg libsrl_b200.so X=0
  run regcarve libsrl_v_regcarve.so X=0
  run regcarvek4 libsrl_v_regcarvek4.so X=0
  run async libsrl_v_async.so X=0
  run asynccarve libsrl_v_asynccarve.so X=0
  run reg_nopdl libsrl_b200.so SRL_PDL=0
done
