#!/bin/bash
# Multi-GPU sweep at N GPUs (N = $1): the 2-rank tests + the peer-exchange check, then cfg2 weak, cfg5 weak, cfg4 strong.
N=${1:-8}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tests/dist_p2p_check.py 2>&1 | grep "p2p exchange" | tee gpurun_out/r2n_p2p_n$N.txt
show() { python - <<PY
import json
try:
    d=json.loads(open('$1').read().strip().splitlines()[-1])
    pc=d['parity_check']; pc.pop('what',None)
    print('$2 N=%d %s step %.1f us value %.4g e2e %.3g (resident %.3g) parity ok=%s table=%s excused=%s' % (d['n_gpus'], d['scaling'], d['ms_per_step']*1e3, d['value'], d['e2e']['value'], d['e2e']['policy_outputs_resident']['value'], pc['ok'], pc.get('stats_table_bit_exact_on_every_rank'), pc.get('boundary_elements_excused')))
except Exception as e:
    print('$2 FAILED', e)
PY
}
run() { # name config scaling extra
  timeout 600 python bench.py --gpus $N --config $2 --scaling $3 --steps 500 --warmup 10 --e2e-steps 20 --no-cpu-baseline --no-extras $4 > gpurun_out/r2n_$1_n$N.json 2> gpurun_out/r2n_$1_n$N.err
  tail -1 gpurun_out/r2n_$1_n$N.err | cut -c1-300; show gpurun_out/r2n_$1_n$N.json $1
}
run cfg2_weak cfg2_atari_large weak
run cfg5_weak cfg5_hns_scale weak
run cfg4_strong cfg4_football_11v11 strong
run cfg2_weak_table cfg2_atari_large weak --no-fuse-stats
