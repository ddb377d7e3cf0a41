#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for n in 1 2; do
timeout 600 python bench.py --gpus $n --config cfg4_football_11v11 --scaling strong --steps 100 --warmup 5 --e2e-steps 5 --no-cpu-baseline --no-extras > gpurun_out/r2j_cfg4_n$n.json 2> gpurun_out/r2j_n$n.err
tail -2 gpurun_out/r2j_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2j_cfg4_n$n.json').read().strip().splitlines()[-1])
pc=d['parity_check']; pc.pop('what')
print('N=$n cfg4 strong step %.1f us value %.3g parity %s' % (d['ms_per_step']*1e3, d['value'], pc))
PY
done
