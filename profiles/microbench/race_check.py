"""Determinism stress: the same eager / graph step N times on the same inputs -- gradients, loss outputs and the statistics
table must be bit-identical every time (any difference is a race).  python profiles/microbench/race_check.py [shuffle_block]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from srl_b200 import ops, synth
from srl_b200.hotpath import HotPath
blk = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = synth.CONFIGS["cfg2_atari_large"]
dev = torch.device("cuda", 0)
s = synth.make_sample_scalars(cfg, seed=0)
E, Mb, T, N = cfg.epochs, cfg.minibatches, cfg.T, cfg.N
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
bad = 0
for trial in range(6):
    hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**bench.hyper_kwargs(cfg)),
                 bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps, epochs=E, minibatches=Mb, seed=0, device=dev,
                 shuffle_block=blk)
    # poison what the scan has to produce, so that a read ahead of its write shows
    hp.lane_aos.fill_(float("nan")); hp.lane_part.fill_(float("nan")); hp.adv.fill_(float("nan")); hp.ret.fill_(float("nan"))
    hp.pack.fill_(float("nan")); hp.perm.fill_(-1)
    hp.load_sample({k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, N))).pin_memory() for k, v in s.items()})
    g = torch.Generator(device="cpu").manual_seed(1)
    pol_all = (torch.randn((E, Mb, 3, T, hp.n_mb), generator=g) * 0.1).to(dev)
    pol_dev = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(Mb)] for e in range(E)]
    ref = None
    for rep in range(40):
        use_graph = rep >= 10
        if rep % 3 == 0:
            flush.zero_()
        if rep % 5 == 0:  # poison again: every step must rewrite what it reads
            hp.lane_aos.fill_(float("nan")); hp.pack.fill_(float("nan")); hp.perm.fill_(-1)
        hp.run_device(pol_dev, use_graph=use_graph)
        hp.step_count = 0
        torch.cuda.synchronize()
        cur = (hp.grads_all.clone(), hp.out.clone())
        hp.ensure_table(); torch.cuda.synchronize()
        tab = hp.global_stats.clone()
        hp._table_valid = False
        if ref is None:
            ref = (cur, tab)
            continue
        dg = not torch.equal(cur[0], ref[0][0]) or not torch.equal(cur[1].nan_to_num(), ref[0][1].nan_to_num())
        dt = not torch.equal(tab, ref[1])
        if dg or dt or not torch.isfinite(cur[0]).all():
            bad += 1
            print(f"trial {trial} rep {rep} graph={use_graph}: grads/out differ={dg} table differs={dt} finite={bool(torch.isfinite(cur[0]).all())}"
                  f" max|dgrad|={(cur[0]-ref[0][0]).abs().max().item():.3e}")
    del hp
print(f"shuffle_block={blk}: {bad} deviating steps of {6*39}")
