#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
timeout 600 python bench.py --steps 500 --warmup 10 --e2e-steps 20 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -3 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2h_bench.json').read().strip().splitlines()[-1]); k=d['kernels']
print('cfg2 step %.1f us K2 %.1f K4 %.1f frac %.3f e2e %.3g (resident %.3g)' % (d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3, d['step']['frac_of_peak'], d['e2e']['value'], d['e2e']['policy_outputs_resident']['value']))
print('trainer_order', d['step_trainer_order']); print('trainer_step', d['trainer_step']); print('extra', d['extra']); print('parity', d['parity_check']); print('clocks', d['clocks']); print('cpu', d['cpu_baseline'])
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err; tail -2 gpurun_out/r2h_ref.err; cut -c1-2500 gpurun_out/r2h_ref.json
