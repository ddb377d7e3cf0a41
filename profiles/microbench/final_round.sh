#!/bin/bash
# End-of-session evidence run (1 GPU): default bench line, reference arm, every config, launch list and one
# `ncu --set full` capture of the default workload.  Outputs under gpurun_out/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r1d}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for c in cfg1_atari_cpu cfg3_smac_27m cfg4_football_11v11 cfg5_hns_scale; do
  python bench.py --config $c --steps 500 --warmup 10 --e2e-steps 10 --no-cpu-baseline > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
kill $SMI
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ppo_loss_kernel|gae_scan|group_stats|philox_perm' -s 8 -c 4 \
  -o gpurun_out/${TAG}_prof_cfg2 -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], d.get("impl", "ours"), "%.4f ms" % d["ms_per_step"], "%.3g" % d["value"], "e2e %.3g" % d["e2e"]["value"],
              "frac", d.get("roofline", {}).get("frac"), "step frac", d.get("step", {}).get("frac_of_peak"))
    except Exception as e:
        print(f, "FAILED", e)
PY
ls -la gpurun_out/${TAG}_*
