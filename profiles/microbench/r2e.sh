#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
run() {
  env SRL_B200_LIB=$PWD/srl_b200/$2 $3 timeout 300 python bench.py --steps 300 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('%-10s step %.1f us K2 %.1f K4 %.1f (warm %.1f) frac %.3f' % ('$1', d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['bytes_per_launch']/k['ppo_loss_kernel']['gbs_l2_warm']/1e3, d['step']['frac_of_peak']))"
}
for rep in 1 2; do
  run default libsrl_b200.so X=0
  run b10 libsrl_v_b10.so X=0
  run d4b6 libsrl_v_d4b6.so X=0
  run d2b10 libsrl_v_d2b10.so X=0
  run t128 libsrl_v_t128.so X=0
done
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'ppo_loss|gae_scan' -s 12 -c 2 \
  -o gpurun_out/r2e_prof_warm -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2e_ncu_warm.log 2>&1
ls -la gpurun_out/r2e_prof*
