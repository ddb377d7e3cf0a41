#!/bin/bash
# Round 2: first run of the pair loss kernel (ppo_loss_pair.cu): GPU suite, then the cfg2 bench.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader; nproc
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for rep in 1 2; do
timeout 300 python bench.py --steps 400 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/r2b.err | tee gpurun_out/r2b_cfg2_$rep.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('cfg2 step %.1f us K2 %.1f K4 %.1f frac %.3f' % (d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, d['step']['frac_of_peak']))"
done
tail -5 gpurun_out/r2b.err
