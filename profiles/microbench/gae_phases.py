"""Per-phase cycle counts of the GAE kernel (library built with -DSRL_DEBUG_PHASES) + event timings."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from srl_b200 import build, synth
dbg = os.path.join(ROOT, "gpurun_out", "libsrl_dbg.so")
extra = [a for a in sys.argv if a.startswith("-D")]
if extra and "--phases" not in sys.argv:
    dbg = os.path.join(ROOT, "gpurun_out", "libsrl_var.so")
    build.build(out=dbg, extra_flags=extra)
    os.environ["SRL_B200_LIB"] = dbg
if "--phases" in sys.argv:
    build.build(out=dbg, extra_flags=["-DSRL_DEBUG_PHASES"])
    os.environ["SRL_B200_LIB"] = dbg
from srl_b200 import ops
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2_atari_large"]
s = synth.make_sample_scalars(cfg, 0)
d = {k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, cfg.N))).cuda() for k, v in s.items()}
adv = torch.empty_like(d["value"]); ret = torch.empty_like(d["value"])
part = torch.empty((8, cfg.N), dtype=torch.float64, device="cuda")
run = lambda: ops.gae_scan(d["reward"], d["value"], d["done"], d["truncated"], d["on_reset"], cfg.gamma, cfg.lmbda, row_lo=0, row_hi=cfg.T, adv=adv, ret=ret, lane_part=part)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run(); torch.cuda.synchronize()
if "--phases" in sys.argv:
    print("-- second launch (L2 warm)", flush=True); run(); torch.cuda.synchronize()
    print("-- after L2 flush", flush=True); flush.zero_(); run(); torch.cuda.synchronize()
    sys.exit(0)
for warm in (True, False):
    ts = []
    for _ in range(30):
        if not warm: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    print(f"{cfg.name} gae_scan {'L2-warm' if warm else 'cold'}: median {ts[len(ts)//2]:.2f} us  min {ts[0]:.2f} us (stream launch, incl. ~3 us launch)")
