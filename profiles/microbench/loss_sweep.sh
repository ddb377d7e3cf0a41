#!/bin/bash
# Builds tuning variants of the library on the GPU box and prints the kernel timings bench.py measures for each.
# usage: profiles/microbench/loss_sweep.sh "<mb>:<unroll> ..."   (SRL_LOSS_MIN_BLOCKS : SRL_LOSS_UNROLL : SRL_LOSS_STAGES)
summ() {
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    k = d["kernels"]
    print(sys.argv[1], "ms/step %.4f" % d["ms_per_step"], "step_frac %.3f" % d["step"]["frac_of_peak"],
          {n: ("%.2f us" % (v["ms_per_launch"] * 1e3), "%d GB/s" % v["gbs"], "warm %d" % v["gbs_l2_warm"]) for n, v in k.items()})
except Exception as e:
    print(sys.argv[1], "ERR", e)
PY
}
mkdir -p gpurun_out
for v in ${1:-"3:1"}; do
  IFS=: read mb un stg <<< "$v"; stg=${stg:-2}   # stages 0 = no cp.async pipeline (register path)
  lib=gpurun_out/libsrl_mb${mb}_u${un}_s${stg}.so
  pipe=1; s2=$stg; if [ "$stg" = "0" ]; then pipe=0; s2=2; fi
  python -m srl_b200.build --out $lib -- -DSRL_LOSS_MIN_BLOCKS=$mb -DSRL_LOSS_UNROLL=$un -DSRL_LOSS_STAGES=$s2 -DSRL_LOSS_PIPE=$pipe $EXTRA_NVCC > /dev/null || continue
  for cfg in ${CFGS:-cfg2_atari_large cfg5_hns_scale}; do
    SRL_B200_LIB=$lib python bench.py --config $cfg --steps 300 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/sweep_${cfg}_${mb}_${un}_${stg}.json 2> gpurun_out/sweep.err
    summ "mb=$mb unroll=$un stages=$stg $cfg" gpurun_out/sweep_${cfg}_${mb}_${un}_${stg}.json
  done
  rm -f $lib
done
