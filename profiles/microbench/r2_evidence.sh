#!/bin/bash
# Evidence run (1 GPU): GPU suite, default bench line + reference arm, every config, launch list and ncu captures of the
# default workload, assembly bench.  Outputs under gpurun_out/ (copied into profiles/ afterwards).  TAG = $1.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2}
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for c in cfg1_atari_cpu cfg3_smac_27m cfg4_football_11v11 cfg5_hns_scale; do
  timeout 600 python bench.py --config $c --steps 500 --warmup 10 --e2e-steps 10 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ppo_loss|gae_scan|group_stats|philox_perm' -s 6 -c 3 \
  -o gpurun_out/${TAG}_prof_cfg2 -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_ncu_full.log 2>&1
for ct in 1 8; do
  timeout 300 python profiles/microbench/assembly_bench.py --copy-threads $ct 2>&1 | tail -1 > gpurun_out/${TAG}_assembly_ct$ct.json
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], d.get("impl", "ours"), "%.4f ms" % d["ms_per_step"], "%.4g" % d["value"], "e2e %.3g" % d["e2e"]["value"],
              "frac", d.get("roofline", {}).get("frac"), "step frac", d.get("step", {}).get("frac_of_peak"), "parity", (d.get("parity_check") or {}).get("ok"))
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("gpurun_out/${TAG}_assembly_ct*.json")):
    print(f, open(f).read()[:400])
PY
