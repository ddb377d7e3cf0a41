#!/bin/bash
# Round 2, first GPU call: the GPU suite on this round's box, the same-box A/B of the K4 build variants left unmeasured in
# round 1 (prebuilt here: srl_b200/libsrl_v_*.so travel with the snapshot), and the assembly bench with srl_host_copy on.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
nproc
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
SRL_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -3
run() {  # tag, library, extra env
  for c in ${CFGS:-cfg2_atari_large}; do
    env SRL_B200_LIB=$PWD/srl_b200/$2 $3 python bench.py --config $c --steps 400 --warmup 10 --e2e-steps 3 --no-cpu-baseline 2> gpurun_out/k4v.err |
      python -c "
import json, sys
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); k = d['kernels']
    print('%-22s %-18s step %.1f us  K2 %.1f  K4 %.1f' % ('$1', '$c', d['ms_per_step'] * 1e3, k['gae_scan_kernel']['ms_per_launch'] * 1e3, k['ppo_loss_kernel']['ms_per_launch'] * 1e3))
except Exception as e:
    print('$1 $c FAILED', e)"
  done
}
for rep in 1 2; do
  run default libsrl_b200.so SRL_X=0
  run lanes2 libsrl_b200.so SRL_LOSS_LANES=2
  run nc_unroll2 libsrl_v_nc2.so SRL_X=0
  run lanes2+nc_unroll2 libsrl_v_nc2.so SRL_LOSS_LANES=2
  run norm_fp32 libsrl_v_norm32.so SRL_X=0
  run lanes2+norm_fp32 libsrl_v_norm32.so SRL_LOSS_LANES=2
  run lanes2+64regs libsrl_v_mb4.so SRL_LOSS_LANES=2
  run shuffle8 libsrl_b200.so SRL_X=0
done
python bench.py --steps 400 --warmup 10 --e2e-steps 3 --no-cpu-baseline --shuffle-block 8 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('shuffle_block=8 step %.1f us K2 %.1f K4 %.1f' % (d['ms_per_step']*1e3, k['gae_scan_kernel']['ms_per_launch']*1e3, k['ppo_loss_kernel']['ms_per_launch']*1e3))"
for ct in 1 4 8; do
  python profiles/microbench/assembly_bench.py --copy-threads $ct 2>&1 | tail -1 | tee gpurun_out/r2a_assembly_ct$ct.json | cut -c1-600
done
