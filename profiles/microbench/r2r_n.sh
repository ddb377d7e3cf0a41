#!/bin/bash
# N-GPU A/B of bench.py argument sets on one box: r2r_n.sh <N> <tag> "<args>" ...
N=$1; TAG=$2; shift; shift
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
i=0
for A in "$@"; do
  i=$((i+1)); [ "$A" = "-" ] && A=""
  timeout 600 python bench.py --gpus $N --steps 500 --warmup 10 --e2e-steps 10 --no-cpu-baseline --no-extras $A > gpurun_out/r2r_${TAG}_${i}_n$N.json 2> gpurun_out/r2r_${TAG}_${i}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2r_${TAG}_${i}_n$N.json').read().strip().splitlines()[-1])
    pc=d['parity_check'] or {}
    print('N=%d [%s] %s step %.2f us value %.4g e2e %.3g (resident %.3g) parity ok=%s table=%s excused=%s' % (d['n_gpus'], '$A', d['scaling'], d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], pc.get('ok'), pc.get('stats_table_bit_exact_on_every_rank'), pc.get('boundary_elements_excused')))
except Exception as e:
    print('[$A] FAILED', e); print(open('gpurun_out/r2r_${TAG}_${i}_n$N.err').read()[-800:])
PY
done
