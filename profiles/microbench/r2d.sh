#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
for rep in 1 2; do
timeout 300 python bench.py --steps 400 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2d.err | tee gpurun_out/r2d_cfg2_$rep.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('cfg2 step %.1f us K2 %.1f K4 %.1f frac %.3f' % (d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, d['step']['frac_of_peak']))"
done
tail -3 gpurun_out/r2d.err
ncu --set full --clock-control none --import-source on -k regex:'ppo_loss' -s 6 -c 1 \
  -o gpurun_out/r2d_prof_pair -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2d_ncu_full.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:'ppo_loss|gae_scan' -s 12 -c 2 \
  -o gpurun_out/r2d_prof_warm -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/r2d_ncu_warm.log 2>&1
ls -la gpurun_out/r2d_prof*
