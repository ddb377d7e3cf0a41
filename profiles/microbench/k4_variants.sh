#!/bin/bash
# Same-box A/B of the loss-kernel candidates that are in the tree but not measured yet (profiles/r1d_notes.md).
# Run in ONE gpurun call (boxes differ by ~2 %):   gpurun --timeout 900 -- 'bash profiles/microbench/k4_variants.sh'
# Variant libraries are built next to the product library (srl_b200/build/ is gpurun-ignored) and removed afterwards.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "## two-lane kernel: bit-equality of the gradients with the four-lane kernel"
SRL_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -3
python -m srl_b200.build --out $PWD/srl_b200/libsrl_v_nc2.so -- -DSRL_LOSS_POLICY_NC=1 -DSRL_LOSS_UNROLL=2 > /dev/null
python -m srl_b200.build --out $PWD/srl_b200/libsrl_v_norm32.so -- -DSRL_LOSS_NORM_FP32=1 > /dev/null
python -m srl_b200.build --out $PWD/srl_b200/libsrl_v_mb4.so -- -DSRL_LOSS_MIN_BLOCKS=4 > /dev/null  # two-lane kernel at 64 registers: 32 warps per SM, 116 B of spills
run() {  # tag, library, extra env
  for c in ${CFGS:-cfg2_atari_large cfg5_hns_scale}; do
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
  CFGS=cfg2_atari_large run lanes2+64regs libsrl_v_mb4.so SRL_LOSS_LANES=2
done
rm -f srl_b200/libsrl_v_nc2.so srl_b200/libsrl_v_norm32.so srl_b200/libsrl_v_mb4.so
