"""GPU parity of the whole step (HotPath) against the oracle's hot_path_ref: GAE once, Philox minibatches,
PopArt update per epoch, fused loss per minibatch -- eager launches, CUDA-graph replay and the host-buffer
entry point must all agree with the CPU restatement of MultiAgentPPO.step's inner loop (mappo.py:240-299)."""
import numpy as np
import pytest
import torch

from oracle import ref_math as M
from srl_b200 import synth
from tests.util import assert_close_ref, assert_grad_close

pytestmark = pytest.mark.gpu


def _hp_kwargs(cfg):
    return dict(eps_clip=cfg.eps_clip, clip_value=cfg.clip_value, dual_clip=cfg.dual_clip, c_clip=cfg.c_clip,
                value_loss=cfg.value_loss, value_loss_weight=cfg.value_loss_weight,
                entropy_bonus_weight=cfg.entropy_bonus_weight,
                value_loss_config=({"delta": cfg.value_loss_delta} if cfg.value_loss == "huber" else None))


def _setup(cfg, seed, fuse_gather=True, **hp_kw):
    from srl_b200 import ops
    from srl_b200.hotpath import HotPath
    s = synth.make_sample_scalars(cfg, seed)
    pol = synth.make_policy_outputs(cfg, s, seed + 1)
    hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**_hp_kwargs(cfg)),
                 bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps, epochs=cfg.epochs,
                 minibatches=cfg.minibatches, seed=77, popart=cfg.popart, popart_beta=0.99, fuse_gather=fuse_gather,
                 **hp_kw)
    hp.load_sample(s)
    # oracle
    batch = {k: torch.from_numpy(v.reshape(cfg.L, cfg.N, 1)).float() for k, v in s.items()}
    batch.update({k: torch.from_numpy(v.reshape(cfg.epochs, cfg.T, cfg.N, 1)) for k, v in pol.items()})
    pa = M.RunningMeanStdRef((1,), beta=0.99) if cfg.popart else None
    ref = M.hot_path_ref(batch, M.LossHyper(**_hp_kwargs(cfg)), cfg.gamma, cfg.lmbda, cfg.epochs, cfg.minibatches, seed=77,
                         popart=pa, bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps,
                         lanes_per_env=cfg.A)
    # policy outputs per (epoch, minibatch) in the oracle's lane order
    n = cfg.N // cfg.minibatches
    pol_dev = []
    for e in range(cfg.epochs):
        if cfg.minibatches > 1:
            env = M.philox_perm_ref(77, e, cfg.B).astype(np.int64)
            perm = (env[:, None] * cfg.A + np.arange(cfg.A)[None]).reshape(-1)
        else:
            perm = np.arange(cfg.N)
        row = []
        for j in range(cfg.minibatches):
            idx = perm[j * n:(j + 1) * n]
            row.append(tuple(torch.from_numpy(np.ascontiguousarray(pol[k][e].reshape(cfg.T, cfg.N)[:, idx])).cuda()
                             for k in ("new_logp", "v_pred", "entropy")))
        pol_dev.append(row)
    return hp, pol_dev, ref, pa


def _compare(hp, cfg, ref, pa):
    torch.cuda.synchronize()
    assert np.array_equal(hp.adv.cpu().numpy(), ref["adv"].numpy()[..., 0]), "adv not bit-exact"
    assert np.array_equal(hp.ret.cpu().numpy(), ref["ret"].numpy()[..., 0]), "ret not bit-exact"
    out = hp.out.cpu().numpy()
    for k, r in enumerate(ref["per_minibatch"]):
        e, j = divmod(k, cfg.minibatches)
        msum = out[k, 9]
        assert_close_ref(out[k, 0], r["loss"], what=f"loss[{e},{j}]")
        assert_close_ref(out[k, 1], r["policy_loss"], what="policy_loss")
        assert_close_ref(out[k, 2], r["value_loss"], what="value_loss")
        for q, name in enumerate(("g_logp", "g_value", "g_entropy")):
            assert_grad_close(hp.grads[e][j][q].cpu(), r[name][..., 0], msum, what=f"{name}[{e},{j}]")
        for name, slot in (("advantage", 4), ("importance_weight", 5), ("clip_ratio", 6), ("value_targets", 7)):
            assert_close_ref(out[k, slot], r["stats"][name], what=name)
    if pa is not None:
        st = hp.popart_state.cpu().numpy()
        np.testing.assert_allclose(st[:3], [pa.mean.item(), pa.mean_sq.item(), pa.debias.item()], rtol=1e-12)
        assert st[3] == cfg.epochs


CASES = {
    "atari_mb": synth.PathConfig("atari_mb", T=16, B=32, epochs=2, minibatches=4, p_end=0.06, clip_value=True,
                                 dual_clip=False, value_loss="huber", value_loss_delta=10.0, value_loss_weight=1.0),
    "smac_popart": synth.PathConfig("smac_popart", T=12, B=8, A=5, epochs=2, minibatches=2, p_end=0.08, lmbda=0.95,
                                    value_loss="huber", value_loss_delta=10.0, popart=True, dead_agent_frac=0.2),
    "boot_burn": synth.PathConfig("boot_burn", T=10, B=12, bootstrap_steps=5, burn_in_steps=3, epochs=1, minibatches=1,
                                  p_end=0.1, popart=True, value_loss="mse"),
    "one_pass": synth.PathConfig("one_pass", T=33, B=50, epochs=1, minibatches=1, p_end=0.05),
    # the pair kernel's corners: an odd first loss row and an odd number of loss rows (half-used row pairs at both ends)
    "mb_burn_odd": synth.PathConfig("mb_burn_odd", T=13, B=24, bootstrap_steps=3, burn_in_steps=3, epochs=2, minibatches=3,
                                    p_end=0.08, clip_value=True, dual_clip=True, value_loss="mse"),
    # several column tiles per minibatch, CTA ranges that cross column and minibatch boundaries
    "mb_wide": synth.PathConfig("mb_wide", T=21, B=1536, epochs=2, minibatches=2, p_end=0.05, clip_value=True,
                                dual_clip=False, value_loss="huber", value_loss_delta=10.0),
    # the scan kernel that computes the permutations and every minibatch's per-CTA shares (256 <= N < ~9500 lanes): the loss
    # kernel adds the shares instead of gathering lane items -- cfg2's width, agents in groups, a half-filled last scan CTA
    "mb_shares": synth.PathConfig("mb_shares", T=6, B=4096, epochs=4, minibatches=8, p_end=0.1, clip_value=True,
                                  value_loss="huber", value_loss_delta=10.0),
    "mb_shares_agents": synth.PathConfig("mb_shares_agents", T=9, B=68, A=4, epochs=3, minibatches=2, p_end=0.1,
                                         dead_agent_frac=0.2),
    # an odd minibatch width: the pack form on the row-tile kernel (one lane per thread)
    "mb_odd_width": synth.PathConfig("mb_odd_width", T=9, B=27, epochs=1, minibatches=3, p_end=0.1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_step_eager_matches_oracle(name):
    cfg = CASES[name]
    hp, pol_dev, ref, pa = _setup(cfg, seed=3)
    hp.run_device(pol_dev, use_graph=False)
    _compare(hp, cfg, ref, pa)
    assert hp.count_launches() > 0


@pytest.mark.parametrize("name", ["atari_mb", "smac_popart"])
def test_step_graph_replay_matches_oracle(name):
    cfg = CASES[name]
    hp, pol_dev, ref, pa = _setup(cfg, seed=4)
    hp.run_device(pol_dev, use_graph=True)  # capture (with a warm-up run) + first replay
    if cfg.popart:  # warm-up and capture runs advanced the PopArt EMA; restart it for the comparison
        hp.popart_state.zero_()
        hp.popart_ms.copy_(torch.tensor([0.0, 0.1, 0.0, 0.1], dtype=torch.float64))
        hp.run_device(pol_dev, use_graph=True)
    _compare(hp, cfg, ref, pa)


@pytest.mark.parametrize("name", ["mb_wide", "mb_shares", "mb_shares_agents"])
def test_step_uses_the_scan_shares_and_agrees_with_the_lane_gather(name, monkeypatch):
    """Where the scan computes the permutations itself the batched loss adds the scan's per-CTA shares of its minibatch
    (HotPath._part_valid); with SRL_MB_PART=0 it gathers the minibatch's lane items as before.  Same sums up to the order
    of the float64 additions: gradients within 1e-12 relative of each other, eager and from the graph."""
    cfg = CASES[name]
    a, pol_dev, ref, pa = _setup(cfg, seed=8)
    monkeypatch.setenv("SRL_MB_PART", "0")
    b, _, _, _ = _setup(cfg, seed=8)
    monkeypatch.delenv("SRL_MB_PART")
    assert a.mb_part is not None and b.mb_part is None
    a.run_device(pol_dev, use_graph=False)
    b.run_device(pol_dev, use_graph=False)
    assert a._part_valid and not b._part_valid
    _compare(a, cfg, ref, pa)
    torch.testing.assert_close(a.grads_all, b.grads_all, rtol=1e-6, atol=1e-12)
    torch.testing.assert_close(a.out, b.out, rtol=1e-9, atol=1e-12)
    eager = a.grads_all.clone()
    a.grads_all.zero_()
    a.run_device(pol_dev, use_graph=True)
    torch.cuda.synchronize()
    assert torch.equal(a.grads_all, eager)


@pytest.mark.parametrize("name", ["atari_mb", "mb_shares", "smac_popart", "one_pass"])
def test_step_recorded_plan_matches_oracle(name):
    """HotPath.run_device(plan=True): the first call records the step's C-ABI calls (an eager step), the following calls
    issue them again as they are.  Both must be the oracle's step; a replay reproduces the recording bit for bit."""
    cfg = CASES[name]
    hp, pol_dev, ref, pa = _setup(cfg, seed=9)
    assert hp.plan_supported()
    hp.run_device(pol_dev, plan=True)  # records
    _compare(hp, cfg, ref, pa)
    first = hp.grads_all.clone()
    assert 1 <= hp.plan_launch_calls <= hp.count_launches()  # C-ABI calls; one call may launch two kernels
    hp.grads_all.zero_()
    hp.out.zero_()
    if cfg.popart:  # the recording advanced the PopArt EMA; restart it for the comparison
        hp.popart_state.zero_()
        hp.popart_ms.copy_(torch.tensor([0.0, 0.1, 0.0, 0.1], dtype=torch.float64))
    hp.run_device(pol_dev, plan=True)  # replays
    _compare(hp, cfg, ref, pa)
    assert torch.equal(hp.grads_all, first)


def test_step_explicit_gather_equals_fused_gather():
    cfg = CASES["atari_mb"]
    a, pol_dev, ref, pa = _setup(cfg, seed=5, fuse_gather=True, batch_losses=False)
    b, _, _, _ = _setup(cfg, seed=5, fuse_gather=False)
    a.run_device(pol_dev, use_graph=False)
    b.run_device(pol_dev, use_graph=False)
    torch.cuda.synchronize()
    for e in range(cfg.epochs):
        for j in range(cfg.minibatches):
            for q in range(3):
                assert torch.equal(a.grads[e][j][q], b.grads[e][j][q])
    assert b.count_launches() == a.count_launches() + cfg.epochs * cfg.minibatches


@pytest.mark.parametrize("name", ["atari_mb", "mb_shares"])
def test_host_entry_point_roundtrip(name):
    """HotPath.run_host: host buffers in and out.  It launches the loss once per EPOCH (each epoch as soon as its policy
    outputs have landed): with the scan's minibatch shares (mb_shares) every such launch starts at its epoch's first table slot."""
    cfg = CASES[name]
    hp, pol_dev, ref, pa = _setup(cfg, seed=6)
    n = cfg.N // cfg.minibatches
    pol_host = torch.stack([torch.stack([torch.stack(list(trip)) for trip in row]) for row in pol_dev]).cpu().pin_memory()
    out_host = dict(adv=torch.empty((cfg.L, cfg.N)).pin_memory(), ret=torch.empty((cfg.L, cfg.N)).pin_memory(),
                    grads=torch.empty((cfg.epochs, cfg.minibatches, 3, cfg.T, n)).pin_memory(),
                    out=torch.empty((cfg.epochs * cfg.minibatches, 16), dtype=torch.float64).pin_memory())
    s = synth.make_sample_scalars(cfg, 6)
    nbytes = hp.run_host(s, pol_host, out_host, use_graph=False)
    ref_grads = out_host["grads"].clone()  # permutations of step 0, the ones pol_host was laid out for
    assert hp._part_valid == (name == "mb_shares")
    assert nbytes["h2d_bytes"] == cfg.L * cfg.N * 15 + cfg.epochs * cfg.T * cfg.N * 12
    assert nbytes["d2h_bytes"] == cfg.L * cfg.N * 8 + cfg.epochs * cfg.T * cfg.N * 12 + cfg.epochs * cfg.minibatches * 128
    assert np.array_equal(out_host["adv"].numpy(), ref["adv"].numpy()[..., 0])
    for k, r in enumerate(ref["per_minibatch"]):
        e, j = divmod(k, cfg.minibatches)
        assert_close_ref(out_host["out"][k, 0], r["loss"], what="loss")
        assert_grad_close(out_host["grads"][e, j, 0], r["g_logp"][..., 0], float(out_host["out"][k, 9]), what="g_logp")
    # float32 flags, as the reference's tests feed them, are narrowed on the host and give the same answer
    s32 = {k: v.astype(np.float32) for k, v in s.items()}
    hp.run_host(s32, pol_host, out_host, use_graph=False)
    assert np.array_equal(out_host["adv"].numpy(), ref["adv"].numpy()[..., 0])
    # the per-epoch-graph variant (arguments frozen at capture: the permutations of step 0) returns the same bytes
    out_host["grads"].zero_()
    hp.run_host(s, pol_host, out_host, use_graph=True)
    hp.run_host(s, pol_host, out_host, use_graph=True)
    assert torch.equal(out_host["grads"], ref_grads)
    assert np.array_equal(out_host["adv"].numpy(), ref["adv"].numpy()[..., 0])


def test_block_shuffle_matches_oracle():
    """shuffle_block = 4: the permutation moves 4-environment blocks (16-byte runs of every float32 leaf), which the
    loss kernel fetches with 128-bit loads; results must equal the oracle run on the same lane order."""
    from srl_b200 import ops
    from srl_b200.hotpath import HotPath
    cfg = CASES["atari_mb"]
    blk = 4
    s = synth.make_sample_scalars(cfg, 8)
    pol = synth.make_policy_outputs(cfg, s, 9)
    hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**_hp_kwargs(cfg)),
                 epochs=cfg.epochs, minibatches=cfg.minibatches, seed=77, shuffle_block=blk)
    hp.load_sample(s)
    batch = {k: torch.from_numpy(v.reshape(cfg.L, cfg.N, 1)).float() for k, v in s.items()}
    batch.update({k: torch.from_numpy(v.reshape(cfg.epochs, cfg.T, cfg.N, 1)) for k, v in pol.items()})
    ref = M.hot_path_ref(batch, M.LossHyper(**_hp_kwargs(cfg)), cfg.gamma, cfg.lmbda, cfg.epochs, cfg.minibatches, seed=77,
                         lanes_per_env=blk)
    n = cfg.N // cfg.minibatches
    pol_dev = []
    for e in range(cfg.epochs):
        env = M.philox_perm_ref(77, e, cfg.B // blk).astype(np.int64)
        perm = (env[:, None] * blk + np.arange(blk)[None]).reshape(-1)
        pol_dev.append([tuple(torch.from_numpy(np.ascontiguousarray(pol[k][e].reshape(cfg.T, cfg.N)[:, perm[j * n:(j + 1) * n]])).cuda()
                              for k in ("new_logp", "v_pred", "entropy")) for j in range(cfg.minibatches)])
    hp.run_device(pol_dev, use_graph=False)
    _compare(hp, cfg, ref, None)


@pytest.mark.parametrize("hp_kw", [dict(use_pack=False), dict(batch_losses=False), dict(use_pack=False, batch_losses=False),
                                   dict(fuse_stats=False), dict(fuse_stats=False, use_pack=False)])
@pytest.mark.parametrize("name", ["atari_mb", "smac_popart"])
def test_step_variants_match_oracle(name, hp_kw):
    """The same step with the loss pack and / or the batched launch switched off (the per-minibatch launches the
    trainer uses) agrees with the oracle too, eagerly and replayed from a CUDA graph."""
    cfg = CASES[name]
    hp, pol_dev, ref, pa = _setup(cfg, 21, **hp_kw)
    hp.run_device(pol_dev, use_graph=False)
    _compare(hp, cfg, ref, pa)
    hp.step_count = 0
    hp.run_device(pol_dev, use_graph=True)  # capture (with a warm-up run) + first replay
    hp.popart_state.zero_()
    hp.popart_ms.copy_(torch.tensor([0.0, 0.1, 0.0, 0.1], dtype=torch.float64))
    hp.step_count = 0
    hp.run_device(pol_dev, use_graph=True)
    _compare(hp, cfg, ref, pa)


def test_full_size_cfg2_pack_and_batch_equal_per_minibatch_launches():
    """BASELINE cfg2 at full size (T=128, B=4096, 4 x 8 minibatches): the batched launch on K2's pack and the 32
    per-minibatch launches on the five gathered leaves give bit-identical gradients and the same loss scalars."""
    from srl_b200 import ops
    from srl_b200.hotpath import HotPath
    cfg = synth.CONFIGS["cfg2_atari_large"]
    s = synth.make_sample_scalars(cfg, 0)
    kw = dict(gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**_hp_kwargs(cfg)), epochs=cfg.epochs,
              minibatches=cfg.minibatches, seed=1)
    a = HotPath(cfg.L, cfg.B, cfg.A, **kw)
    b = HotPath(cfg.L, cfg.B, cfg.A, use_pack=False, batch_losses=False, **kw)
    g = torch.Generator(device="cuda").manual_seed(3)
    n = cfg.N // cfg.minibatches
    pol = [[(-torch.rand(cfg.T, n, device="cuda", generator=g), torch.randn(cfg.T, n, device="cuda", generator=g),
             torch.rand(cfg.T, n, device="cuda", generator=g)) for _ in range(cfg.minibatches)] for _ in range(cfg.epochs)]
    for hp in (a, b):
        hp.load_sample(s)
        hp.run_device(pol, use_graph=False)
    torch.cuda.synchronize()
    assert torch.equal(a.adv, b.adv) and torch.equal(a.ret, b.ret) and torch.equal(a.perm, b.perm)
    assert torch.equal(a.grads_all, b.grads_all), "pack / batched gradients differ from the per-minibatch launches"
    # the masked sums add up to 32 terms in fp32 before they enter the float64 accumulators, and the two launch shapes
    # group the terms differently: equal to ~1e-6, an order of magnitude inside the 1e-5 contract
    np.testing.assert_allclose(a.out.cpu().numpy(), b.out.cpu().numpy(), rtol=2e-6, atol=1e-9)
    # round trip property at full size: the same step from a CUDA graph reproduces itself
    a.run_device(pol, use_graph=True)
    a.step_count = 0
    ref_grads = a.grads_all.clone()
    a.run_device(pol, use_graph=True)
    torch.cuda.synchronize()
    assert torch.equal(a.grads_all, ref_grads)


def _slot_state_words(ws):
    """What must be zero between launches: the ticket (bytes 0-3), the published statistics (bytes 32-63) and every
    partial row (bytes 64-); bytes 4-31 are the deferred mode's hand-over to srl_ppo_loss_finalize (n_rows, sum(mask),
    weights) and keep their last values."""
    w = ws.view(torch.int32).view(ws.shape[0], -1)
    return int(w[:, 0].count_nonzero()) + int(w[:, 8:].count_nonzero())


def test_workspace_slots_are_left_all_zero_and_steps_repeat_bit_for_bit():
    """The loss workspace contract (include/srl_b200.h): every kernel leaves a slot's state zero -- the pair kernel's immediate
    mode reads "zero" as "not there yet" for the published statistics and the partial rows -- whichever path used it last:
    the batched launch, the per-minibatch deferred launches + srl_ppo_loss_finalize, and again the batched launch."""
    import numpy as np
    from srl_b200 import ops, synth
    from srl_b200.hotpath import HotPath
    cfg = synth.PathConfig("ws", T=16, B=256, epochs=2, minibatches=4, clip_value=True)
    s = synth.make_sample_scalars(cfg, 3)
    hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(clip_value=True), epochs=2,
                 minibatches=4, seed=0)
    hp.load_sample({k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, cfg.N))) for k, v in s.items()})
    g = torch.Generator().manual_seed(0)
    pol_all = (torch.randn((2, 4, 3, cfg.T, hp.n_mb), generator=g) * 0.1).cuda()
    pol = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(4)] for e in range(2)]
    assert hp.fuse_stats and hp.pack is not None  # the pair kernel with its own statistics
    hp.run_device(pol, use_graph=False)
    torch.cuda.synchronize()
    assert _slot_state_words(hp.workspace) == 0
    first = (hp.grads_all.clone(), hp.out.clone())
    hp.step_count = 0
    hp.run_trainer_order(pol, use_graph=False)  # deferred launches + finalize on the same slots
    torch.cuda.synchronize()
    assert _slot_state_words(hp.workspace) == 0
    # (the table adds the lanes in another order than the loss kernel's own statistics: last-bit differences)
    torch.testing.assert_close(hp.grads_all, first[0], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(hp.out[:, :10], first[1][:, :10], rtol=1e-6, atol=1e-9)
    hp.step_count = 0
    hp.run_device(pol, use_graph=False)
    torch.cuda.synchronize()
    assert _slot_state_words(hp.workspace) == 0
    assert torch.equal(hp.grads_all, first[0]) and torch.equal(hp.out, first[1])


def test_scan_first_and_permutation_first_orders_agree(monkeypatch):
    """K2 -> K5a -> K4 (the permutation kernel beside the scan, completing behind it) and round 1's K5a -> K2 -> K4 produce
    the same step, bit for bit, eagerly and from a graph."""
    import numpy as np
    from srl_b200 import ops, synth
    from srl_b200.hotpath import HotPath
    cfg = synth.PathConfig("ord", T=32, B=512, epochs=2, minibatches=4, clip_value=True)
    s = synth.make_sample_scalars(cfg, 5)
    g = torch.Generator().manual_seed(1)
    res = {}
    for order in ("0", "1"):
        monkeypatch.setenv("SRL_PERM_FIRST", order)
        hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(clip_value=True), epochs=2,
                     minibatches=4, seed=7)
        hp.perm.fill_(-1)
        hp.load_sample({k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, cfg.N))) for k, v in s.items()})
        g.manual_seed(1)
        pol_all = (torch.randn((2, 4, 3, cfg.T, hp.n_mb), generator=g) * 0.1).cuda()
        pol = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(4)] for e in range(2)]
        for use_graph in (False, True, True):
            hp.run_device(pol, use_graph=use_graph)
            hp.step_count = 0
            torch.cuda.synchronize()
            cur = (hp.perm.clone(), hp.adv.clone(), hp.grads_all.clone(), hp.out.clone())
            for a, b in zip(cur, res.setdefault("ref", cur)):
                assert torch.equal(a, b)
