#!/bin/bash
# Round 2: pair loss kernel -- thread-count / occupancy variants on one box, then one ncu --set full capture with sources.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() {
  env SRL_B200_LIB=$PWD/srl_b200/$2 $3 timeout 300 python bench.py --steps 300 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('%-10s step %.1f us K2 %.1f K4 %.1f (warm %.1f) frac %.3f' % ('$1', d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['bytes_per_launch']/k['ppo_loss_kernel']['gbs_l2_warm']/1e3, d['step']['frac_of_peak']))"
}
for rep in 1 2; do
  run default libsrl_b200.so X=0
  run t128b3 libsrl_v_t128b3.so X=0
  run t256b2 libsrl_v_t256b2.so X=0
  run t256b1 libsrl_v_t256b1.so X=0
  run t64b8 libsrl_v_t64b8.so X=0
done
ncu --set full --clock-control none --import-source on -k regex:'ppo_loss' -s 6 -c 2 \
  -o gpurun_out/r2c_prof_pair -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2c_ncu_full.log 2>&1
ls -la gpurun_out/r2c_prof_pair*
