#!/bin/bash
# 2 GPUs: the multi-rank tests, then cfg2 weak at N=2 and N=1 on the same box.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_trainer.py -m gpu -q -x 2>&1 | tail -8
for n in 1 2; do
timeout 600 python bench.py --gpus $n --steps 500 --warmup 10 --e2e-steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2i_cfg2_n$n.json 2> gpurun_out/r2i_n$n.err
tail -2 gpurun_out/r2i_n$n.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2i_cfg2_n$n.json').read().strip().splitlines()[-1])
print('N=$n cfg2 step %.1f us value %.3g e2e %.3g (resident %.3g) parity %s' % (d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], d['parity_check']))
PY
done
