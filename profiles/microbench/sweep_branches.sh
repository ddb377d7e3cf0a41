#!/bin/bash
# usage: sweep_branches.sh  -> ms/step of the cfg2 bench for several CUDA-graph branch counts
for b in 1 4 8 16 32; do
  python bench.py --steps 300 --warmup 10 --no-cpu-baseline --branches $b --e2e-steps 3 $EXTRA 2>/dev/null > /tmp/b.json
  python - "$b" <<'PY'
import json, sys
d = json.loads(open("/tmp/b.json").read().strip().splitlines()[-1])
print("branches", sys.argv[1], "ms_per_step", round(d["ms_per_step"], 4), "l2_warm", round(d["step"]["ms_per_step_l2_warm"], 4))
PY
done
