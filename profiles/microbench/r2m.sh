#!/bin/bash
# 2 GPUs: multi-rank tests, then cfg2 weak at N=2 with the in-kernel exchange (and without: --no-fuse-stats), N=1 beside it.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
show() { python - <<PY
import json
d=json.loads(open('$1').read().strip().splitlines()[-1])
pc=d['parity_check']; pc.pop('what',None)
print('$2 step %.1f us value %.3g e2e %.3g (resident %.3g) parity %s' % (d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], pc))
PY
}
timeout 600 python bench.py --gpus 1 --steps 500 --warmup 10 --e2e-steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2m_cfg2_n1.json 2> gpurun_out/r2m_n1.err; tail -2 gpurun_out/r2m_n1.err; show gpurun_out/r2m_cfg2_n1.json "N=1 cfg2"
timeout 600 python bench.py --gpus 2 --steps 500 --warmup 10 --e2e-steps 20 --no-cpu-baseline --no-extras > gpurun_out/r2m_cfg2_n2.json 2> gpurun_out/r2m_n2.err; tail -2 gpurun_out/r2m_n2.err; show gpurun_out/r2m_cfg2_n2.json "N=2 cfg2 in-kernel exchange"
timeout 600 python bench.py --gpus 2 --steps 500 --warmup 10 --e2e-steps 20 --no-cpu-baseline --no-extras --no-fuse-stats > gpurun_out/r2m_cfg2_n2_table.json 2> gpurun_out/r2m_n2t.err; tail -2 gpurun_out/r2m_n2t.err; show gpurun_out/r2m_cfg2_n2_table.json "N=2 cfg2 table exchange"
