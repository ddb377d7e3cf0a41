#!/bin/bash
# Last evidence run of round 2 (1 GPU).  First a same-box A/B of the scan's one open knob (-DSRL_WS_SKIP_PAD=1: the first
# chunk's padding-only chain left out); the rest runs with the winner (gpurun_out/<tag>_winner.txt says which -- the product
# default is set to it afterwards, same source + same flag = same code): GPU suite, smoke, default bench line + reference arm,
# every config, launch list and ncu captures of the default workload.  Outputs under gpurun_out/ (copied into profiles/).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TAG=${1:-r2x}
ab() {
  SRL_B200_LIB=$PWD/srl_b200/$2 timeout 120 python bench.py --steps 500 --warmup 10 --e2e-steps 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ab_$1.json 2> gpurun_out/${TAG}_ab_$1.err
}
ab base1 libsrl_b200.so; ab skip1 libsrl_b200_skip0.so; ab base2 libsrl_b200.so; ab skip2 libsrl_b200_skip0.so
WIN=$(python - <<PY
import json
def us(n):
    d = json.loads(open('gpurun_out/${TAG}_ab_%s.json' % n).read().strip().splitlines()[-1])
    assert d['parity_check']['ok']
    return d['ms_per_step'] * 1e3
try:
    b, s = (us('base1') + us('base2')) / 2, (us('skip1') + us('skip2')) / 2
    print('libsrl_b200_skip0.so' if s < b - 0.12 else 'libsrl_b200.so')
    open('gpurun_out/${TAG}_winner.txt', 'w').write('base %.3f us, skip0 %.3f us (two runs each)\n' % (b, s))
except Exception as e:
    print('libsrl_b200.so')
    open('gpurun_out/${TAG}_winner.txt', 'w').write('A/B failed: %r\n' % (e,))
PY
)
echo "winner: $WIN ($(cat gpurun_out/${TAG}_winner.txt))" | tee -a gpurun_out/${TAG}_winner.txt
export SRL_B200_LIB=$PWD/srl_b200/$WIN
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for c in cfg1_atari_cpu cfg3_smac_27m cfg4_football_11v11 cfg5_hns_scale; do
  timeout 120 python bench.py --config $c --steps 500 --warmup 10 --e2e-steps 10 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_$c.json 2> gpurun_out/${TAG}_bench_$c.err
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_cfg2.csv \
  python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'ppo_loss|gae_scan|group_stats|philox_perm' -s 6 -c 3 \
  -o gpurun_out/${TAG}_prof_cfg2 -f python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/${TAG}_ncu_full.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("bench_")[1], d.get("impl", "ours"), "%.4f ms" % d["ms_per_step"], "%.4g" % d["value"], "e2e %.3g" % d["e2e"]["value"],
              "frac", d.get("roofline", {}).get("frac"), "step frac", d.get("step", {}).get("frac_of_peak"), "parity", (d.get("parity_check") or {}).get("ok"))
    except Exception as e:
        print(f, "FAILED", e)
PY
