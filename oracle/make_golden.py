"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    python -O -m oracle.make_golden      # -O as in production (apps/main.py:44): the data-invariant asserts of
                                         # gae.py:67-78 do not hold for V-trace-scaled deltas at truncations

Every fixture stores the exact inputs and what the reference's own functions returned for them
(loaded through oracle/ref_loader.py): `MultiAgentPPO._compute_adv_and_value_target`,
`MultiAgentPPO._compute_loss` + `loss.backward()`, `modules.masked_normalization`,
`modules.PopArtValueHead`, `base.namedarray.recursive_aggregate`.  The GPU box has no reference;
tests there compare the CUDA path and the oracle restatement against these files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from srl_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

LOSS_VARIANTS = {
    # name: (trainer kwargs, popart, shape cfg)
    "atari": dict(kw=dict(eps_clip=0.2, clip_value=True, dual_clip=False, value_loss="huber",
                          value_loss_config=dict(delta=10.0), value_loss_weight=1.0, entropy_bonus_weight=0.01),
                  popart=False, A=1),
    "smac": dict(kw=dict(eps_clip=0.2, clip_value=False, dual_clip=True, c_clip=3, value_loss="huber",
                         value_loss_config=dict(delta=10.0), value_loss_weight=1.0, entropy_bonus_weight=0.01),
                 popart=True, A=3, dead=0.2),
    "football": dict(kw=dict(eps_clip=0.2, clip_value=True, dual_clip=True, c_clip=3, value_loss="huber",
                             value_loss_config=dict(delta=10.0), value_loss_weight=1.0, normalize_old_value=True),
                     popart=True, A=2),
    "hns_mse": dict(kw=dict(eps_clip=0.2, clip_value=False, dual_clip=False, value_loss="mse", value_loss_weight=0.5),
                    popart=True, A=1),
    "mse_clip_dual": dict(kw=dict(eps_clip=0.1, clip_value=True, value_eps_clip=0.05, dual_clip=True, c_clip=2.0,
                                  value_loss="mse", value_loss_weight=0.5, entropy_bonus_weight=0.02),
                          popart=False, A=1),
    "smoothl1": dict(kw=dict(eps_clip=0.3, clip_value=True, dual_clip=True, value_loss="smoothl1",
                             value_loss_config=dict(beta=0.5), value_loss_weight=0.7),
                     popart=False, A=1),
    "huber_default": dict(kw=dict(clip_value=False, dual_clip=True, value_loss="huber"), popart=False, A=1),
}


def _popart_head(R, seed, n_updates=3):
    """A PopArtValueHead whose running stats have seen a few batches (so mean/std are non-trivial)."""
    g = torch.Generator().manual_seed(seed)
    head = R.popart.PopArtValueHead(4, 1, beta=0.99)
    for _ in range(n_updates):
        head.update(torch.randn(50, 1, generator=g) * 3.0 + 1.5, mask=None)
    return head


def _rms_state(head):
    rms = head._PopArtValueHead__rms
    return np.array([rms._RunningMeanStd__mean.item(), rms._RunningMeanStd__mean_sq.item(),
                     rms._RunningMeanStd__debiasing_term.item()], dtype=np.float64)


def _sample_namedarray(R, t):
    NA = R.namedarray.NamedArray
    return R.trainer.SampleBatch(obs=None, on_reset=t["on_reset"], done=t["done"], truncated=t["truncated"],
                                 reward=t["reward"],
                                 analyzed_result=NA(value=t["value"], log_probs=t["old_logp"], adv=t.get("adv"),
                                                    ret=t.get("ret")))


def gen_gae(R):
    out = {}
    for name, cfg, popart, vtrace in (
        ("cfg1", synth.CONFIGS["cfg1_atari_cpu"], False, False),
        ("cfg1_boot50", synth.PathConfig("cfg1_boot50", T=80, B=32, bootstrap_steps=50, p_end=0.02), False, False),
        ("smac_small", synth.PathConfig("smac_small", T=40, B=6, A=5, p_end=0.05, gamma=0.99, lmbda=0.95), True, False),
        ("vtrace", synth.PathConfig("vtrace", T=33, B=20, p_end=0.05, gamma=0.97, lmbda=0.9), False, True),
        ("ragged", synth.PathConfig("ragged", T=17, B=37, p_end=0.1, gamma=0.9, lmbda=0.8), True, False),
    ):
        s = synth.make_sample_scalars(cfg, seed=11)
        t = {k: torch.from_numpy(v).float() for k, v in s.items()}  # the prefetcher's .float() (api/trainer.py:217)
        head = _popart_head(R, 5) if popart else None
        tr = R.mappo.MultiAgentPPO(ref_loader.FakePolicy(popart_head=head), discount_rate=cfg.gamma,
                                   gae_lambda=cfg.lmbda, popart=popart, vtrace=vtrace)
        analyzed = None
        extra = {}
        if vtrace:
            g = torch.Generator().manual_seed(3)
            newlp = t["old_logp"][:-1] + 0.3 * torch.randn(t["old_logp"][:-1].shape, generator=g)
            analyzed = R.mappo.SampleAnalyzedResult(old_action_log_probs=t["old_logp"][:-1], new_action_log_probs=newlp,
                                                    state_values=None)
            extra["vtrace_new_logp"] = newlp.numpy()
        adv, ret = tr._compute_adv_and_value_target(_sample_namedarray(R, t), analyzed)
        fx = dict(s)
        fx.update(extra)
        fx.update(adv=adv.numpy(), ret=ret.numpy(), gamma=np.float64(cfg.gamma), lmbda=np.float64(cfg.lmbda),
                  popart=np.bool_(popart), vtrace=np.bool_(vtrace))
        if popart:
            m, sd = head._PopArtValueHead__rms.mean_std()
            fx.update(popart_mean_std=np.array([m.item(), sd.item()], dtype=np.float64))
        out[name] = fx
    for name, fx in out.items():
        np.savez_compressed(os.path.join(GOLDEN, f"gae_{name}.npz"), **fx)
    return list(out)


def gen_loss(R):
    names = []
    for name, v in LOSS_VARIANTS.items():
        A = v["A"]
        cfg = synth.PathConfig(name, T=12, B=10, A=A, p_end=0.08, gamma=0.99, lmbda=0.95, popart=v["popart"],
                               dead_agent_frac=v.get("dead", 0.0), num_actions=(18,))
        s = synth.make_sample_scalars(cfg, seed=21)
        pol = synth.make_policy_outputs(cfg, s, seed=22, epochs=1)
        t = {k: torch.from_numpy(x).float() for k, x in s.items()}
        head = _popart_head(R, 9) if v["popart"] else None
        fake = ref_loader.FakePolicy(popart_head=head,
                                     denormalize_value_during_rollout=v["kw"].get("normalize_old_value", False))
        tr = R.mappo.MultiAgentPPO(fake, discount_rate=cfg.gamma, gae_lambda=cfg.lmbda, popart=v["popart"], **v["kw"])
        adv, ret = tr._compute_adv_and_value_target(_sample_namedarray(R, t), None)
        pad = lambda x: torch.cat([x, torch.zeros_like(x[:1])], 0)  # mappo.py:254-256
        t["adv"], t["ret"] = pad(adv), pad(ret)
        L = cfg.L
        lo, hi = 0, L - 1
        sample = _sample_namedarray(R, t)
        valid = sample[lo:hi]
        mask = 1 - t["on_reset"][lo + 1:hi + 1]  # mappo.py:260-261
        pre_state = _rms_state(head) if head is not None else None
        if v["popart"]:
            fake.update_popart(valid.analyzed_result.ret, mask=mask)  # mappo.py:263-264
        nl = torch.from_numpy(pol["new_logp"][0]).requires_grad_(True)
        vp = torch.from_numpy(pol["v_pred"][0]).requires_grad_(True)
        en = torch.from_numpy(pol["entropy"][0]).requires_grad_(True)
        analyzed = R.mappo.SampleAnalyzedResult(old_action_log_probs=t["old_logp"][lo:hi], new_action_log_probs=nl,
                                                state_values=vp, entropy=en)
        loss, res = tr._compute_loss(valid, analyzed, mask)
        loss.backward()
        stats = {f"stat_{k}": np.float64(getattr(res, k).detach().float().mean().item())
                 for k in ("advantage", "entropy", "policy_loss", "value_loss", "importance_weight", "clip_ratio",
                           "value_targets")}
        if res.denorm_value is not None:
            stats["stat_denorm_value"] = np.float64(res.denorm_value.mean().item())
        fx = dict(s)
        fx.update(new_logp=pol["new_logp"][0], v_pred=pol["v_pred"][0], entropy=pol["entropy"][0],
                  adv=t["adv"].numpy(), ret=t["ret"].numpy(), loss=np.float64(loss.item()), g_logp=nl.grad.numpy(),
                  g_value=vp.grad.numpy(), g_entropy=en.grad.numpy(), gamma=np.float64(cfg.gamma),
                  lmbda=np.float64(cfg.lmbda), **stats)
        if head is not None:
            m, sd = head._PopArtValueHead__rms.mean_std()
            fx.update(popart_state_before=pre_state, popart_state_after=_rms_state(head),
                      popart_mean_std_after=np.array([m.item(), sd.item()], dtype=np.float64), popart_beta=np.float64(0.99))
        np.savez_compressed(os.path.join(GOLDEN, f"loss_{name}.npz"), **fx)
        names.append(name)
    return names


def gen_masknorm_popart(R):
    rng = np.random.Generator(np.random.PCG64(31))
    adv = rng.standard_normal((10, 8, 1)).astype(np.float32)
    mask = rng.integers(0, 2, (10, 8, 1)).astype(np.float32)
    out = R.utils.masked_normalization(torch.from_numpy(adv), torch.from_numpy(mask)).numpy()
    out_nomask = R.utils.masked_normalization(torch.from_numpy(adv), None).numpy()
    # PopArt sequence: updates with and without mask, normalise / denormalise after each
    head = R.popart.PopArtValueHead(4, 1, beta=0.999)
    xs = rng.standard_normal((4, 30, 1)).astype(np.float32) * 2.0 + 0.5
    ms = (rng.random((4, 30, 1)) < 0.7).astype(np.float32)
    y = rng.standard_normal((20, 1)).astype(np.float32)
    states, norms, denorms = [], [], []
    for i in range(4):
        head.update(torch.from_numpy(xs[i]), mask=torch.from_numpy(ms[i]) if i % 2 else None)
        states.append(_rms_state(head))
        norms.append(head.normalize(torch.from_numpy(y)).numpy())
        denorms.append(head.denormalize(torch.from_numpy(y)).numpy())
    np.savez_compressed(os.path.join(GOLDEN, "masknorm_popart.npz"), adv=adv, mask=mask, norm_adv=out,
                        norm_adv_nomask=out_nomask, pa_x=xs, pa_mask=ms, pa_y=y, pa_states=np.stack(states),
                        pa_norm=np.stack(norms), pa_denorm=np.stack(denorms), pa_beta=np.float64(0.999))


def gen_stack(R):
    """recursive_aggregate(np.stack(axis=1)) incl. a leaf that is None in some samples (namedarray.py:588-633)."""
    NA = R.namedarray.NamedArray
    rng = np.random.Generator(np.random.PCG64(41))
    L, B = 5, 6
    samples, flat = [], {}
    for b in range(B):
        obs = NA(frame=rng.integers(0, 256, (L, 2, 3, 3), dtype=np.uint8), vec=rng.standard_normal((L, 7)).astype(np.float32))
        reward = rng.standard_normal((L, 1)).astype(np.float32)
        on_reset = rng.integers(0, 2, (L, 1), dtype=np.uint8)
        trunc = None if b % 2 == 0 else rng.integers(0, 2, (L, 1), dtype=np.uint8)
        samples.append(R.trainer.SampleBatch(obs=obs, reward=reward, on_reset=on_reset, truncated=trunc))
        flat[f"s{b}.obs.frame"], flat[f"s{b}.obs.vec"] = obs.frame, obs.vec
        flat[f"s{b}.reward"], flat[f"s{b}.on_reset"] = reward, on_reset
        if trunc is not None:
            flat[f"s{b}.truncated"] = trunc
    agg = R.namedarray.recursive_aggregate(samples, lambda x: np.stack(x, axis=1))
    for k, v in R.namedarray.flatten(agg):
        if v is not None:
            flat[f"out.{k}"] = v
    flat["out_keys"] = np.array([k for k, v in R.namedarray.flatten(agg) if v is not None])
    np.savez_compressed(os.path.join(GOLDEN, "stack.npz"), **flat)


def gen_nstep(R):
    """modules.n_step_return (legacy/algorithm/modules/n_step_return.py:11-50): random trajectories with episode ends
    and truncations, several n; plus the reference's own known-answer inputs (legacy/tests/modules_test.py:180-209)."""
    g = torch.Generator().manual_seed(11)
    out = {}
    for name, (rows, B, n, gamma) in dict(a=(40, 7, 1, 0.99), b=(40, 7, 3, 0.99), c=(65, 33, 5, 0.997), d=(12, 4, 12, 0.9)).items():
        reward = torch.randn(rows, B, 1, generator=g)
        nex_value = torch.randn(rows, B, 1, generator=g) * 3
        nex_done = (torch.rand(rows, B, 1, generator=g) < 0.08).float()
        nex_trunc = ((torch.rand(rows, B, 1, generator=g) < 0.08).float() * (1 - nex_done))
        ret = R.nstep.n_step_return(n, reward, nex_value, nex_done, nex_trunc, gamma)
        out.update({f"{name}.reward": reward.numpy(), f"{name}.nex_value": nex_value.numpy(),
                    f"{name}.nex_done": nex_done.numpy().astype(np.uint8), f"{name}.nex_truncated": nex_trunc.numpy().astype(np.uint8),
                    f"{name}.n": np.int64(n), f"{name}.gamma": np.float64(gamma), f"{name}.ret": ret.numpy()})
    np.savez_compressed(os.path.join(GOLDEN, "nstep.npz"), **out)
    return sorted({k.split(".")[0] for k in out})


def _episodes(seed, dtype, W, n_eps):
    """Random finished episodes: lengths 2..40 (and one of length 1), final step done / truncated / without value."""
    rng = np.random.Generator(np.random.PCG64(seed))
    eps = []
    for k in range(n_eps):
        n = 1 if k == 3 else int(rng.integers(2, 41))
        kind = k % 3  # 0: truncated with final value, 1: done with final value, 2: done without analyzed_result
        eps.append(dict(reward=rng.standard_normal((n, W)).astype(dtype), value=(3 * rng.standard_normal((n, W))).astype(dtype),
                        truncated=np.full((W,), kind == 0, dtype=np.uint8), has_value=kind != 2))
    return eps


def gen_traj_gae(R):
    """TrajGAE.process (legacy/algorithm/modules/gae.py:100-139) of the unmodified reference on random episodes, fed as
    lists of per-step SampleBatches exactly as an actor worker does (actor_worker.py:152-155); plus the two hand-made
    episodes of the reference's own test (legacy/tests/modules_test.py:140-178)."""
    SB, AR = R.trainer.SampleBatch, R.namedarray.NamedArray
    out = {}
    for name, (dtype, W, gamma, lmbda) in dict(f32w1=(np.float32, 1, 0.99, 0.97), f32w3=(np.float32, 3, 0.997, 0.95),
                                               f64w2=(np.float64, 2, 0.9, 0.8)).items():
        eps = _episodes(7, dtype, W, 12)
        proc = R.gae.TrajGAE(gamma=gamma, lmbda=lmbda)
        rw, vl, tr, hv, lens, adv, ret = [], [], [], [], [], [], []
        for ep in eps:
            n = ep["reward"].shape[0]
            memory = []
            for i in range(n):
                last = i == n - 1
                ar = None if (last and not ep["has_value"]) else AR(value=ep["value"][i].copy(), adv=None, ret=None)
                trunc = ep["truncated"] if last else np.zeros((W,), np.uint8)
                memory.append(SB(obs=None, reward=ep["reward"][i].copy(), analyzed_result=ar,
                                 done=(1 - trunc).astype(np.uint8) if last else np.zeros((W,), np.uint8), truncated=trunc))
            memory = proc.process(memory)
            a = np.stack([m.analyzed_result.adv for m in memory[:-1]]) if n > 1 else np.zeros((0, W), dtype)
            r = np.stack([m.analyzed_result.ret for m in memory[:-1]]) if n > 1 else np.zeros((0, W), dtype)
            assert a.dtype == dtype and r.dtype == dtype
            rw.append(ep["reward"]); vl.append(ep["value"]); tr.append(ep["truncated"]); hv.append(ep["has_value"])
            lens.append(n); adv.append(a); ret.append(r)
        out.update({f"{name}.reward": np.concatenate(rw), f"{name}.value": np.concatenate(vl),
                    f"{name}.final_truncated": np.stack(tr), f"{name}.final_has_value": np.array(hv, np.uint8),
                    f"{name}.lens": np.array(lens, np.int64), f"{name}.adv": np.concatenate(adv),
                    f"{name}.ret": np.concatenate(ret), f"{name}.gamma": np.float64(gamma), f"{name}.lmbda": np.float64(lmbda)})
    np.savez_compressed(os.path.join(GOLDEN, "traj_gae.npz"), **out)
    return sorted({k.split(".")[0] for k in out})


def gen_gae_general(R):
    """modules.gae_trace itself (gae.py:8-97) on what MultiAgentPPO never passes: vector critics, per-element gamma /
    lambda tensors, a ready-made importance ratio.  Run with -O: the data asserts of gae.py:69-77 assume zero rewards
    and values at episode ends, which random vector-critic data does not satisfy."""
    g = torch.Generator().manual_seed(23)
    out = {}
    for name, (T, B, Nc, gt, lt, vt) in dict(nc3=(21, 9, 3, False, False, False), gam=(33, 17, 1, True, False, False),
                                             lam=(18, 40, 1, False, True, False), both_nc2=(25, 6, 2, True, True, False),
                                             all=(30, 11, 4, True, True, True), vt_nc1=(16, 35, 1, False, False, True)).items():
        cfg = synth.PathConfig(name, T=T, B=B, p_end=0.08)
        s = synth.make_sample_scalars(cfg, seed=5)
        flags = {k: torch.from_numpy(s[k]).float() for k in ("done", "truncated", "on_reset")}
        reward = torch.randn(T, B, Nc, generator=g) * (1 - flags["on_reset"][1:])
        value = torch.randn(T + 1, B, Nc, generator=g) * 2
        gamma = (0.9 + 0.1 * torch.rand(T, B, 1, generator=g)) if gt else 0.99
        lmbda = (0.8 + 0.2 * torch.rand(T, B, 1, generator=g)) if lt else 0.95
        ratio = torch.exp(0.5 * torch.randn(T, B, 1, generator=g)) if vt else None
        adv = R.gae.gae_trace(reward, value, flags["truncated"], flags["done"], flags["on_reset"], gamma, lmbda,
                              vtrace=vt, imp_ratio=ratio, rho=1.0, c=0.9)
        out.update({f"{name}.reward": reward.numpy(), f"{name}.value": value.numpy(), f"{name}.adv": adv.numpy(),
                    **{f"{name}.{k}": s[k] for k in ("done", "truncated", "on_reset")}})
        out[f"{name}.gamma"] = gamma.numpy() if gt else np.float64(gamma)
        out[f"{name}.lmbda"] = lmbda.numpy() if lt else np.float64(lmbda)
        if vt:
            out[f"{name}.imp_ratio"] = ratio.numpy()
    np.savez_compressed(os.path.join(GOLDEN, "gae_general.npz"), **out)
    return sorted({k.split(".")[0] for k in out})


def gen_wire(R):
    """base.namedarray.dumps of the unmodified reference for the methods that need no third-party codec
    ('raw_bytes', 'pickle_dict'; blosc is not installed) on a nested sample with a None leaf, a bool leaf and metadata.
    Stored as one uint8 stream + frame lengths per method, plus the leaves for comparison."""
    NA = R.namedarray.NamedArray
    rng = np.random.Generator(np.random.PCG64(17))
    L = 5
    x = R.trainer.SampleBatch(
        obs=NA(frame=rng.integers(0, 255, (L, 2, 6, 6), dtype=np.uint8), vec=rng.standard_normal((L, 3)).astype(np.float32)),
        on_reset=(rng.random((L, 1)) < 0.3), done=np.zeros((L, 1), np.uint8), truncated=None,
        action=NA(x=rng.integers(0, 18, (L, 1)).astype(np.int32)), reward=rng.standard_normal((L, 1)).astype(np.float32),
        analyzed_result=NA(value=rng.standard_normal((L, 1)).astype(np.float32), log_probs=rng.standard_normal((L, 1)),
                           adv=None, ret=None),
        policy_state=NA(hx=rng.standard_normal((L, 1, 4)).astype(np.float32)),
        policy_version_steps=np.full((L, 1), 7, np.int64), sampling_weight=2.5)
    out = {}
    # the compressed methods need `blosc` (absent): the unmodified reference runs with a stand-in codec injected under
    # that name (tests/util.py::StandInBlosc), which pins the FRAMING of these methods, not the codec
    from tests.util import StandInBlosc
    sys.modules["blosc"] = StandInBlosc
    for method in ("raw_bytes", "pickle_dict", "raw_compress", "compress_pickle", "obs_compress",
                   "compress_except_policy_state"):
        frames = R.namedarray.dumps(x, method=method)
        out[f"{method}.stream"] = np.frombuffer(b"".join(frames), dtype=np.uint8)
        out[f"{method}.lens"] = np.array([len(f) for f in frames], np.int64)
        back = R.namedarray.loads(frames)
        assert back.metadata == x.metadata
    del sys.modules["blosc"]
    for k, v in R.namedarray.flatten(x):
        if v is not None:
            out[f"leaf.{k}"] = v
    out["none_leaves"] = np.array([k for k, v in R.namedarray.flatten(x) if v is None])
    np.savez_compressed(os.path.join(GOLDEN, "wire.npz"), **out)
    return sorted(k for k in out if k.startswith("leaf."))


# ---------------------------------------------------------------------------------------------------------------------
# A10: the step orchestration.  The UNMODIFIED MultiAgentPPO.step (mappo.py:219-328) driven on CPU through the host
# stand-in of the prefetcher (ref_loader._HostPrefetch keeps the one-call delay), with a tiny policy (tests/doubles.py;
# its PopArt head is the reference's own class here).  What these fixtures pin: the prefetch delay, the epoch loop,
# tail_len (:243), GAE before PopArt before the loss (:249-266), the optimizer step per epoch, stats averaged over the
# epochs (:293-303), entropy-coefficient decay (:310-311), the info aggregation (:317-324), frames, the version, and the
# adv / ret written back into the host sample (:254-257).
# ---------------------------------------------------------------------------------------------------------------------
TRAINER_OBS_DIM, TRAINER_ACTIONS = 6, 5
TRAINER_CASES = {
    "plain": (dict(T=12, B=16, p_end=0.08),
              dict(clip_value=True, dual_clip=False, value_loss="huber", value_loss_config=dict(delta=10.0),
                   value_loss_weight=1.0, ppo_epochs=2, entropy_decay_per_steps=2, entropy_bonus_decay=0.5)),
    "popart": (dict(T=10, B=6, A=3, p_end=0.1, lmbda=0.95),
               dict(gae_lambda=0.95, dual_clip=True, value_loss="huber", value_loss_config=dict(delta=10.0), popart=True,
                    ppo_epochs=2, max_grad_norm=0.5)),
    "vtrace": (dict(T=9, B=8, p_end=0.1), dict(vtrace=True, dual_clip=False)),
    "boot_burn": (dict(T=9, B=8, bootstrap_steps=3, burn_in_steps=2, p_end=0.1),
                  dict(bootstrap_steps=3, burn_in_steps=2, value_loss="smoothl1", popart=True, ppo_epochs=3,
                       recompute_adv_among_epochs=True)),
}


def trainer_sample_arrays(shape_kw, seed):
    """The leaves of one synthetic sample of a trainer fixture (also what the tests rebuild their samples from)."""
    cfg = synth.PathConfig("trainer_fixture", **shape_kw)
    s = synth.make_sample_scalars(cfg, seed)
    rng = np.random.default_rng(seed + 99)
    lead = s["value"].shape[:-1]
    s["obs_vec"] = rng.standard_normal(lead + (TRAINER_OBS_DIM,)).astype(np.float32)
    s["action_x"] = rng.integers(0, TRAINER_ACTIONS, lead + (1,)).astype(np.float32)
    s["info_episode_return"] = rng.standard_normal(lead + (1,)).astype(np.float32)
    s["info_mask"] = (rng.random(lead + (1,)) < 0.1).astype(np.float32)
    return cfg, s


def gen_trainer(R):
    import json
    from tests.doubles import TinyActorCriticPolicy
    NA = R.namedarray.NamedArray
    n_calls = 4  # the first call only primes the prefetcher
    written = []
    for name, (shape_kw, kw) in TRAINER_CASES.items():
        cfg = synth.PathConfig("trainer_fixture", **shape_kw)
        kw = dict(kw, discount_rate=cfg.gamma, optimizer="sgd", optimizer_config=dict(lr=0.05))
        pol = TinyActorCriticPolicy(TRAINER_OBS_DIM, TRAINER_ACTIONS, device="cpu", popart=kw.get("popart", False), seed=3,
                                    popart_head_cls=R.popart.PopArtValueHead)
        out = {"kwargs_json": np.array(json.dumps(kw)), "shape_json": np.array(json.dumps(shape_kw)),
               "n_calls": np.array(n_calls)}
        for k, v in pol.net.state_dict().items():
            out[f"init/{k}"] = v.detach().cpu().numpy().copy()
        tr = R.mappo.MultiAgentPPO(pol, **kw)
        samples = []
        for it in range(n_calls):
            _, a = trainer_sample_arrays(shape_kw, seed=10 + it)
            for k, v in a.items():
                out[f"call{it}/in/{k}"] = v.copy()
            sb = R.trainer.SampleBatch(obs=NA(vec=a["obs_vec"]), on_reset=a["on_reset"], done=a["done"],
                                       truncated=a["truncated"], action=NA(x=a["action_x"]), reward=a["reward"],
                                       info=NA(episode_return=a["info_episode_return"]), info_mask=a["info_mask"],
                                       analyzed_result=NA(value=a["value"], log_probs=a["old_logp"], adv=None, ret=None))
            samples.append(sb)
            res = tr.step(sb)
            out[f"call{it}/step"] = np.array(res.step)
            out[f"call{it}/stat_keys"] = np.array(sorted(res.stats))
            for k, v in res.stats.items():
                out[f"call{it}/stat/{k}"] = np.array(float(v), dtype=np.float64)
            if it >= 1:  # this call trained on the PREVIOUS sample: its host copy now carries adv / ret (or None)
                prev = samples[it - 1].analyzed_result
                out[f"call{it}/adv_is_none"] = np.array(prev.adv is None)
                if prev.adv is not None:
                    out[f"call{it}/adv"], out[f"call{it}/ret"] = prev.adv.copy(), prev.ret.copy()
            out[f"call{it}/entropy_bonus_weight"] = np.array(float(tr.entropy_bonus_weight))
        for k, v in pol.net.state_dict().items():
            out[f"final/{k}"] = v.detach().cpu().numpy().copy()
        np.savez_compressed(os.path.join(GOLDEN, f"trainer_{name}.npz"), **out)
        written.append(name)
    return written


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    R = ref_loader.load()
    if len(sys.argv) > 1 and sys.argv[1] == "nstep":  # only the n-step fixture (added after the others)
        print("nstep:", gen_nstep(R))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "family":  # TrajGAE + general gae_trace fixtures (added later still)
        print("traj_gae:", gen_traj_gae(R))
        print("gae_general:", gen_gae_general(R))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "wire":
        print("wire:", gen_wire(R))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "trainer":  # the step orchestration (added in round 2)
        print("trainer:", gen_trainer(R))
        return
    print("gae:", gen_gae(R))
    print("loss:", gen_loss(R))
    gen_masknorm_popart(R)
    gen_stack(R)
    print("nstep:", gen_nstep(R))
    print("traj_gae:", gen_traj_gae(R))
    print("gae_general:", gen_gae_general(R))
    print("wire:", gen_wire(R))
    print("trainer:", gen_trainer(R))
    total = sum(os.path.getsize(os.path.join(GOLDEN, f)) for f in os.listdir(GOLDEN))
    print(f"wrote {len(os.listdir(GOLDEN))} files, {total / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
