"""Per-phase cycles (with --phases: library built with -DSRL_DEBUG_PHASES) and launch durations of the loss kernel."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from srl_b200 import build, synth
extra = [a for a in sys.argv if a.startswith("-D")]
if "--phases" in sys.argv:
    extra.append("-DSRL_DEBUG_PHASES")
if extra:
    dbg = os.path.join(ROOT, "gpurun_out", "libsrl_var.so")
    subprocess.check_call([build.find_nvcc()] + build.NVCC_FLAGS + extra + ["-I", build.INCLUDE, "-o", dbg] + [os.path.join(build.CSRC, s) for s in build.SOURCES])
    os.environ["SRL_B200_LIB"] = dbg
from srl_b200 import ops
cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg2_atari_large"]
mode = "gather" if "--gather" in sys.argv else ("full" if "--full" in sys.argv else "contig")
T, N = cfg.T, cfg.N
n = N if mode == "full" else N // max(cfg.minibatches, 1)
rng = np.random.default_rng(0)
f = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32)).cuda()
olp, ov, ret, adv = f(cfg.L, N), f(cfg.L, N), f(cfg.L, N), f(cfg.L, N)
rs = torch.from_numpy((rng.random((cfg.L, N)) < 0.01).astype(np.uint8)).cuda()
nl, vp, en = olp[:T, :n].contiguous() + 0.1 * f(T, n), f(T, n), f(T, n)
idx = torch.randperm(N, device="cuda")[:n].to(torch.int32).contiguous() if mode == "gather" else None
stats = torch.tensor([float(T * n), 0.0, float(T * n), 0, 0, 0, 0, 0], dtype=torch.float64, device="cuda")
hp = ops.LossHyper(clip_value=cfg.clip_value, dual_clip=cfg.dual_clip, value_loss=cfg.value_loss,
                   value_loss_config=({"delta": 10.0} if cfg.value_loss == "huber" else None))
ws = ops.new_loss_workspace("cuda", slots=1)
grads = tuple(torch.empty_like(nl) for _ in range(3))
if mode == "gather":
    smp = (olp[:T], ov[:T], ret[:T], adv[:T], rs[1:T + 1])
else:
    c = lambda x: x[:, :n].contiguous()
    smp = (c(olp)[:T], c(ov)[:T], c(ret)[:T], c(adv)[:T], c(rs)[1:T + 1])
run = lambda: ops.ppo_loss_fwd_bwd(nl, vp, en, *smp, stats, hp, lane_idx=idx, grads=grads, workspace=ws[0], defer=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
run(); torch.cuda.synchronize()
if "--phases" in sys.argv:
    print("-- second launch (L2 warm)", flush=True); run(); torch.cuda.synchronize()
    print("-- after L2 flush", flush=True); flush.zero_(); run(); torch.cuda.synchronize()
    sys.exit(0)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(16): run()
for cold in (False, True):
    ts = []
    for _ in range(20):
        if cold: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / 16)
    ts.sort()
    by = (41 if cfg.clip_value else 37) * T * n
    print(f"{cfg.name} loss [{T}x{n}] {mode}: {'cold first touch' if cold else 'L2-warm'}: {ts[len(ts)//2]:.2f} us/launch "
          f"(graph of 16 serial launches) -> {by / ts[len(ts)//2] / 1e3:.0f} GB/s algorithmic")
