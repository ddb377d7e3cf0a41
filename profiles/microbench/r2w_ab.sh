#!/bin/bash
# Same-box A/B of the warp-specialised scan's tuning knobs (gae_scan_ws.cu): r2w_ab.sh <tag> [steps]
#   base      : product build
#   skip      : -DSRL_WS_SKIP_PAD=1   the scanner leaves out the chain steps of the rows behind the trajectory's end (15 of 144 at L = 129)
#   skipearly : + -DSRL_WS_EARLY_TMA=1  first chunk requested before the other barriers are initialised
#   skipr32   : skip + -DSRL_WS_ROWS=32  chunks of 32 rows (half the hand-offs between the workers and the scanner)
#   all       : the three together
cd "$(dirname "$0")/../.."
TAG=${1:-r2w}; STEPS=${2:-500}
SRL_B200_LIB=$PWD/srl_b200/libsrl_b200_all.so timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_hotpath.py tests/test_gpu_family.py -x -q 2>&1 | tail -2
run() {
  local n=$1 lib=$2
  SRL_B200_LIB=$PWD/srl_b200/$lib timeout 120 python bench.py --steps $STEPS --warmup 10 --e2e-steps 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_${n}.json 2> gpurun_out/${TAG}_${n}.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_${n}.json').read().strip().splitlines()[-1])
    k=d['kernels']
    print('%-10s: step %.2f us (warm %.2f) frac %.3f | K4 %.2f us K2 %.2f us (warm %.2f) | parity %s' % (
      '$n', d['ms_per_step']*1e3, d['step']['ms_per_step_l2_warm']*1e3, d['step']['frac_of_peak'], k['ppo_loss_kernel']['ms_per_launch']*1e3,
      k['gae_scan_kernel']['ms_per_launch']*1e3, k['gae_scan_kernel']['bytes_per_launch']/k['gae_scan_kernel']['gbs_l2_warm']/1e3, d['parity_check']['ok']))
except Exception as ex:
    print('$n: FAILED', ex)
PY
}
run base1 libsrl_b200.so
run skip libsrl_b200_skip.so
run skipearly libsrl_b200_skipearly.so
run skipr32 libsrl_b200_skipr32.so
run all libsrl_b200_all.so
run base2 libsrl_b200.so
