"""How much of the cold-L2 step is the launch itself?  CUDA-event time of tiny graphs right behind the bench's L2 flush
(cold) and back to back (warm): a 1-CTA torch fill, the permutation kernel alone, the scan alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from srl_b200 import ops, synth
cfg = synth.CONFIGS["cfg2_atari_large"]
dev = torch.device("cuda", 0)
s = synth.make_sample_scalars(cfg, 0)
d = {k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, cfg.N))).to(dev) for k, v in s.items()}
adv = torch.empty_like(d["value"]); ret = torch.empty_like(d["value"])
part = torch.empty((8, cfg.N), dtype=torch.float64, device=dev)
perm = torch.empty((cfg.epochs, cfg.N), dtype=torch.int32, device=dev)
tiny = torch.zeros(32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
def flush_l2():
    flush.zero_(); flush_rd.max()
def graph_of(fn):
    st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st): fn()
    torch.cuda.current_stream().wait_stream(st); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g
ops.set_pdl(False)
cases = {
    "torch fill of 32 floats (1 CTA)": lambda: tiny.fill_(1.0),
    "K5a philox_perm alone": lambda: ops.philox_perm(0, 0, cfg.B, cfg.A, out=perm, n_epochs=cfg.epochs),
    "K2 gae_scan alone": lambda: ops.gae_scan(d["reward"], d["value"], d["done"], d["truncated"], d["on_reset"], cfg.gamma, cfg.lmbda,
                                              row_lo=0, row_hi=cfg.T, adv=adv, ret=ret, lane_part=part),
    "nothing (two events back to back)": None,
}
for name, fn in cases.items():
    g = graph_of(fn) if fn else None
    for cold in (True, False):
        ts = []
        for _ in range(60):
            if cold: flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if g: g.replay()
            b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[10:])
        print(f"{name:36s} {'cold (behind the L2 flush)' if cold else 'warm (back to back)       '}: median {ts[len(ts)//2]:6.2f} us  min {ts[0]:6.2f}")

# ---- the same kernels as plain stream launches (the CPU is far ahead of the GPU here: the flush takes > 100 us)
print("-- stream launches instead of a graph")
for name, fn in cases.items():
    if fn is None: continue
    for cold in (True,):
        ts = []
        for _ in range(60):
            if cold: flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        ts = sorted(ts[10:])
        print(f"{name:36s} stream launch, cold: median {ts[len(ts)//2]:6.2f} us  min {ts[0]:6.2f}")

# ---- the whole step: graph against eager launches
sys.path.insert(0, ROOT)
import bench
from srl_b200.hotpath import HotPath
ops.set_pdl(True)
pol = synth.make_policy_outputs(cfg, s, seed=1)
E, Mb, T, N = cfg.epochs, cfg.minibatches, cfg.T, cfg.N
hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**bench.hyper_kwargs(cfg)),
             bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps, epochs=E, minibatches=Mb, seed=0, device=dev)
n = hp.n_mb
hp.load_sample({k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, N))).pin_memory() for k, v in s.items()})
pol_all = torch.randn((E, Mb, 3, T, n), dtype=torch.float32, device=dev) * 0.1
pol_dev = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(Mb)] for e in range(E)]
for use_graph in (True, False):
    ts = []
    for _ in range(200):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); hp.run_device(pol_dev, use_graph=use_graph); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[20:])
    print(f"whole cfg2 step, {'one CUDA graph' if use_graph else 'three stream launches'}, cold: median {ts[len(ts)//2]:6.2f} us  min {ts[0]:6.2f}  mean {sum(ts)/len(ts):6.2f}")
