"""Device-side timeline of the cfg2 step graph (K5a -> K2 -> K4): library built with -DSRL_TIMELINE, every CTA stamps
%globaltimer at its phase boundaries, the host prints when each phase starts / ends relative to the step's first stamp.

    python profiles/microbench/timeline.py [cfg] [-DFLAG ...] [--tag NAME]

Run on the GPU box (builds the instrumented library into gpurun_out/).  The numbers explain where the step's time goes;
they are not bench values (the stamps cost a few stores per CTA).
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from srl_b200 import build, synth

extra = [a for a in sys.argv[1:] if a.startswith("-D")]
tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "tl"
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
lib = os.path.join(ROOT, "gpurun_out", f"libsrl_{tag}.so")
if int(os.environ.get("LOCAL_RANK", "0")) == 0:
    build.build(out=lib, extra_flags=["-DSRL_TIMELINE"] + extra)
else:
    import time
    while not os.path.exists(lib + ".done"):
        time.sleep(0.5)
if int(os.environ.get("LOCAL_RANK", "0")) == 0:
    open(lib + ".done", "w").close()
os.environ["SRL_B200_LIB"] = lib
from srl_b200 import ops
from srl_b200.hotpath import HotPath

names = [a for a in sys.argv[1:] if not a.startswith("-") and a != tag]
cfg = synth.CONFIGS[names[0] if names else "cfg2_atari_large"]
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
pg = None
if world > 1:  # torchrun: the same picture with the statistics exchanged between the ranks inside the loss kernel
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
    pg = dist.group.WORLD
h = ops._lib.load_library()
KC, KS = 2048, 16
buf = torch.zeros((3, KC, KS, 2), dtype=torch.int64, device=dev)
for fn in ("srl_tl_set_perm", "srl_tl_set_gae", "srl_tl_set_loss"):
    f = getattr(h, fn)
    f.argtypes = [ctypes.c_void_p]
    assert f(buf.data_ptr()) == 0, fn

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (hyper_kwargs)

s = synth.make_sample_scalars(cfg, seed=0)
pol = synth.make_policy_outputs(cfg, s, seed=1)
E, Mb, T, N = cfg.epochs, cfg.minibatches, cfg.T, cfg.N
hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**bench.hyper_kwargs(cfg)),
             bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps, epochs=E, minibatches=Mb, seed=0,
             popart=cfg.popart, device=dev, process_group=pg, exchange_timeout_s=30.0)
n = hp.n_mb
hp.load_sample({k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, N))).pin_memory() for k, v in s.items()})
perm = [ops.philox_perm(0, e, cfg.B, cfg.A).long() if Mb > 1 else None for e in range(E)]
pol_all = torch.empty((E, Mb, 3, T, n), dtype=torch.float32, device=dev)
for e in range(E):
    for q, k in enumerate(("new_logp", "v_pred", "entropy")):
        full = torch.from_numpy(pol[k][e].reshape(T, N)).to(dev)
        for j in range(Mb):
            pol_all[e, j, q] = full if Mb == 1 else full.index_select(1, perm[e][j * n:(j + 1) * n])
pol_dev = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(Mb)] for e in range(E)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)


def flush_l2():
    flush.zero_()
    flush_rd.max()


for _ in range(5):
    flush_l2()
    hp.run_device(pol_dev, use_graph=True)
torch.cuda.synchronize()

K_NAMES = {0: "K5a perm", 1: "K2 scan", 2: "K4 loss"}
SLOTS = {
    0: ["entry", "end"],
    1: ["entry", "barriers ready", "all TMA issued", "first chunk landed", "scan done", "last store pass done",
        "lane sums written", "after wait for K5a"],
    2: ["entry", "after griddepcontrol.wait", "statistics ready", "loop done", "row written", "problem finalised",
        "first CTA: local sums ready", "first CTA: sums of all ranks ready"],
}


_one = torch.ones(1, dtype=torch.float64, device=dev)
_sum = torch.zeros(1, dtype=torch.float64, device=dev)


def one(cold=True):
    buf.zero_()
    if cold:
        flush_l2()
    if world > 1:  # line the ranks up on the device, as bench.py does before every timed step
        hp.peer.allreduce_sum(_one, _sum)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    hp.run_device(pol_dev, use_graph=True)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3, buf.cpu().numpy()


ENTRY = {0: 0, 1: 1, 2: 2}  # kernel whose slot 0 is the entry stamp of the same CTA
GHZ = 1.965


def report(us, rec, label):
    """Only entry stamps carry the global timer (256 ns ticks); every stamp carries the SM's cycle counter.  A stamp's place
    on the global axis = its CTA's entry (global) + cycles since that entry / clock."""
    gt = rec[..., 0].astype(np.float64)
    clk = rec[..., 1].astype(np.float64)
    t0 = gt[gt > 0].min()
    print(f"== {label}: step {us:.2f} us by CUDA events; us after the first stamp of the step | cycles after the CTA's entry")
    rows = {}
    for k in sorted(SLOTS):
        e_gt, e_clk = gt[ENTRY[k], :, 0], clk[ENTRY[k], :, 0]
        for sl, nm in enumerate(SLOTS[k]):
            ok = (clk[k, :, sl] > 0) & (e_clk > 0)
            if ok.sum() == 0:
                continue
            cyc = clk[k, :, sl][ok] - e_clk[ok]
            v = (e_gt[ok] - t0) / 1e3 + cyc / GHZ / 1e3
            rows[f"{K_NAMES[k]}: {nm}"] = dict(ctas=int(v.size), min=float(v.min()), median=float(np.median(v)),
                                                max=float(v.max()), cycles_median=float(np.median(cyc)))
            print(f"  {K_NAMES[k]:30s} {nm:28s} ctas {v.size:5d}  us min {v.min():6.2f} med {np.median(v):6.2f} max {v.max():6.2f}"
                  f" | cyc min {cyc.min():7.0f} med {np.median(cyc):7.0f} max {cyc.max():7.0f}")
    return rows


if rank != 0:
    sys.stdout = open(os.devnull, "w")
res = {}
for label, cold in (("cold L2", True), ("cold L2 (again)", True), ("warm L2", False)):
    us, rec = one(cold)
    res[label] = dict(step_us=us, rows=report(us, rec, label))
# globaltimer resolution: distinct consecutive values among all stamps
g = np.unique(rec[..., 0][rec[..., 0] > 0])
if g.size > 2:
    print("globaltimer smallest increment seen: %d ns" % int(np.diff(g).min()))
if rank == 0:
    np.save(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}_{cfg.name}.npy"), rec)
if rank == 0:
    with open(os.path.join(ROOT, "gpurun_out", f"timeline_{tag}_{cfg.name}.json"), "w") as f:
        json.dump(res, f, indent=1)
if world > 1:
    del hp
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
