#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
SRL_B200_LIB=$PWD/srl_b200/libsrl_v_l1t128b8.so timeout 600 python -m pytest tests/test_gpu_hotpath.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4
run() {
  env SRL_B200_LIB=$PWD/srl_b200/$2 $3 timeout 300 python bench.py --steps 300 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2g.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('%-12s step %.1f us K2 %.1f K4 %.1f (warm %.1f) frac %.3f' % ('$1', d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['bytes_per_launch']/k['ppo_loss_kernel']['gbs_l2_warm']/1e3, d['step']['frac_of_peak']))"
}
for rep in 1 2; do
  run reg libsrl_b200.so X=0
  run l1t64b16 libsrl_v_l1t64b16.so X=0
  run l1t128b8 libsrl_v_l1t128b8.so X=0
  run l1t128b6 libsrl_v_l1t128b6.so X=0
  run l1t128b5 libsrl_v_l1t128b5.so X=0
  run l1t64b10 libsrl_v_l1t64b10.so X=0
done
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'ppo_loss' -s 6 -c 1 \
  -o gpurun_out/r2g_prof_l1 -f env SRL_B200_LIB=$PWD/srl_b200/libsrl_v_l1t128b8.so python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
ls -la gpurun_out/r2g_prof*
