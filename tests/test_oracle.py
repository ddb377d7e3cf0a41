"""CPU tests: pin the oracle restatement (oracle/ref_math.py) to the reference.

Three sources of truth, in order of authority:
  1. the reference's own known-answer tests (legacy/tests/modules_test.py), restated here;
  2. fixtures produced by the UNMODIFIED reference (tests/golden/*.npz, made by oracle/make_golden.py);
  3. where /root/reference is present (build container only): the live reference on fresh random inputs.
"""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from oracle import ref_math as M
from srl_b200 import synth
from tests.util import assert_close_ref, load_golden

GAE_FIXTURES = ["cfg1", "cfg1_boot50", "smac_small", "vtrace", "ragged"]
LOSS_FIXTURES = ["atari", "smac", "football", "hns_mse", "mse_clip_dual", "smoothl1", "huber_default"]


def _f32(d, *keys):
    return [torch.from_numpy(d[k]).float() for k in keys]


def _popart_from(mean_std):
    """An oracle RunningMeanStd whose mean_std() returns exactly (mu, sigma)."""
    pa = M.RunningMeanStdRef((1,), beta=0.99)
    mu, sd = float(mean_std[0]), float(mean_std[1])
    pa.debias = torch.ones(1, dtype=torch.float64)
    pa.mean = torch.tensor([mu], dtype=torch.float64)
    pa.mean_sq = torch.tensor([sd * sd + mu * mu], dtype=torch.float64)
    return pa


# ---- 1. reference known-answer tests --------------------------------------------------------------
def test_gae_truncated_known_answer():
    """legacy/tests/modules_test.py:119-138 (hand-computed, gamma = lambda = 0.1)."""
    on_reset = np.array([0, 0, 0, 1, 0, 0, 1, 0, 0], dtype=np.float32)
    rew = np.array([1, 2, 0, 1, 3, 0, 1, 2, 3], dtype=np.float32)
    value = np.array([2, 0, 1, 2, 2, 0, 1, 1, 1], dtype=np.float32)
    truncated = np.array([0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=np.float32)
    done = np.array([0, 0, 0, 0, 0, 1, 0, 0, 0], dtype=np.float32)
    adv = M.gae_trace_ref(torch.from_numpy(rew)[:-1], torch.from_numpy(value), torch.from_numpy(truncated),
                          torch.from_numpy(done), torch.from_numpy(on_reset), 0.1, 0.1).numpy()
    expect = np.array([2.1 * 0.01 - 1, 2.1, 0, -0.8 + 0.01, 1, 0, 0.111, 1.1])
    np.testing.assert_array_almost_equal(adv * (1 - on_reset[1:]), expect * (1 - on_reset[1:]))


def test_gae_vs_independent_numpy():
    """legacy/tests/modules_test.py:91-117: compare with the tianshou-style recursion, abs < 1e-5."""
    rng = np.random.default_rng(0)
    value = rng.standard_normal((101, 8, 1))
    rew = rng.standard_normal((100, 8, 1))
    done = rng.integers(0, 2, (101, 8, 1)).astype(np.float64)
    on_reset = np.concatenate([np.zeros_like(done[:1]), done[:-1]], axis=0)
    rew = rew * (1 - on_reset[1:])
    value = value * (1 - done)
    delta = rew + value[1:] * 0.99 * (1 - done[:-1]) - value[:-1]
    m = (1.0 - done[:-1]) * (0.99 * 0.97)
    gae, expect = 0.0, np.zeros_like(rew)
    for i in range(99, -1, -1):
        gae = delta[i] + m[i] * gae
        expect[i] = gae
    got = M.gae_trace_ref(*(torch.from_numpy(x) for x in (rew, value, np.zeros_like(done), done, on_reset)), 0.99, 0.97)
    assert np.abs(got.numpy() - expect).max() < 1e-5


def test_traj_gae_known_answer():
    """legacy/tests/modules_test.py:140-178 (TrajGAE postprocessor)."""
    adv, _ = M.traj_gae_ref([1, 2, 0], [2, 0, 1], last_truncated=True, last_has_value=True, gamma=0.1, lmbda=0.1)
    np.testing.assert_allclose(adv, [2.1 * 0.01 - 1, 2.1])
    adv, _ = M.traj_gae_ref([1, 3, 0], [2, 2, 0], last_truncated=False, last_has_value=True, gamma=0.1, lmbda=0.1)
    np.testing.assert_allclose(adv, [-0.8 + 0.01, 1])


def test_mask_norm_vs_nanmean():
    """legacy/tests/modules_test.py:38-46,267-271."""
    rng = np.random.default_rng(1)
    adv = rng.standard_normal((10, 8, 1))
    mask = rng.integers(0, 2, (10, 8, 1)).astype(np.float64)
    a = adv.copy()
    a[mask == 0] = np.nan
    expect = (adv - np.nanmean(a)) / (np.nanstd(a) + 1e-5)
    got = M.masked_normalization_ref(torch.from_numpy(adv), torch.from_numpy(mask)).numpy()
    np.testing.assert_almost_equal(got * mask, expect * mask, decimal=6)


def test_mask_norm_two_rank_emulation():
    """legacy/tests/modules_test.py:273-299: halves normalised with summed statistics equal the whole."""
    rng = np.random.default_rng(2)
    adv = torch.from_numpy(rng.standard_normal((1, 16, 1)))
    mask = torch.from_numpy(rng.integers(0, 2, (1, 16, 1)).astype(np.float64))
    whole = M.masked_normalization_ref(adv, mask)
    s = [M.masked_sums_ref(adv[:, h], mask[:, h]) for h in (slice(0, 8), slice(8, 16))]
    tot = tuple(a + b for a, b in zip(*s))
    halves = torch.cat([M.masked_normalization_ref(adv[:, h], mask[:, h], global_sums=tot)
                        for h in (slice(0, 8), slice(8, 16))], dim=1)
    np.testing.assert_almost_equal(halves.numpy(), whole.numpy(), decimal=6)


def test_popart_closed_form():
    """legacy/tests/modules_test.py:301-323: debiased EMA expectations."""
    g = torch.Generator().manual_seed(0)
    pa = M.RunningMeanStdRef((3,), beta=0.999)
    x = torch.randn(100, 3, generator=g)
    x_mean, x_std = x.mean(0), (x.square().mean(0) - x.mean(0).square()).sqrt()
    y = torch.randn(20, 3, generator=g)
    for _ in range(5):
        pa.update(x)
        torch.testing.assert_close(pa.normalize(y), (y - x_mean) / x_std, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(pa.denormalize(y), y * x_std + x_mean, rtol=1e-4, atol=1e-5)
    x2 = torch.randn(20, 3, generator=g)
    mean = x_mean + (x2.mean(0) - x_mean) / (1 - 0.999**6) * 0.001
    mean_sq = x.square().mean(0) + (x2.square().mean(0) - x.square().mean(0)) / (1 - 0.999**6) * 0.001
    pa.update(x2)
    torch.testing.assert_close(pa.normalize(y), (y - mean) / (mean_sq - mean**2).sqrt(), rtol=1e-4, atol=1e-5)


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (counter x4, key x2 -> output x4)."""
    kat = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, out in kat:
        got = M.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert [int(v) for v in got] == out


@pytest.mark.parametrize("n", [1, 2, 3, 5, 32, 100, 512, 4096, 5000])
def test_philox_perm_is_permutation(n):
    p0 = M.philox_perm_ref(1234, 0, n)
    assert p0.dtype == np.int32 and sorted(p0.tolist()) == list(range(n))
    assert np.array_equal(p0, M.philox_perm_ref(1234, 0, n))  # stateless / reproducible
    if n >= 32:
        assert not np.array_equal(p0, M.philox_perm_ref(1234, 1, n))  # epochs differ
        assert not np.array_equal(p0, M.philox_perm_ref(1235, 0, n))  # seeds differ
        assert not np.array_equal(p0, np.arange(n))


def test_philox_perm_mixes():
    """Sanity on shuffle quality: each minibatch of a 4096-lane batch draws from the whole lane range."""
    p = M.philox_perm_ref(7, 3, 4096).reshape(8, 512)
    assert (np.abs(p.mean(axis=1) - 2047.5) < 150).all()
    assert np.abs(np.corrcoef(np.arange(4096), p.reshape(-1))[0, 1]) < 0.05


# ---- 2. fixtures generated by the unmodified reference ----------------------------------------------
@pytest.mark.parametrize("name", GAE_FIXTURES)
def test_gae_matches_reference_fixture(name):
    d = load_golden(f"gae_{name}.npz")
    rew, val, tr, dn, rs, olp = _f32(d, "reward", "value", "truncated", "done", "on_reset", "old_logp")
    pa = _popart_from(d["popart_mean_std"]) if bool(d["popart"]) else None
    kw = {}
    if bool(d["vtrace"]):
        kw = dict(vtrace=True, new_logp=torch.from_numpy(d["vtrace_new_logp"]), old_logp=olp[:-1])
    adv, ret = M.adv_and_value_target_ref(rew, val, tr, dn, rs, float(d["gamma"]), float(d["lmbda"]), popart=pa, **kw)
    # same torch ops in the same order on the same machine: expect (near) bit equality
    assert_close_ref(adv, d["adv"], tol=1e-6, what=f"{name} adv")
    assert_close_ref(ret, d["ret"], tol=1e-6, what=f"{name} ret")


def _loss_hyper(name):
    from oracle.make_golden import LOSS_VARIANTS
    kw = dict(LOSS_VARIANTS[name]["kw"])
    return M.LossHyper(**kw), LOSS_VARIANTS[name]["popart"]


@pytest.mark.parametrize("name", LOSS_FIXTURES)
def test_loss_matches_reference_fixture(name):
    d = load_golden(f"loss_{name}.npz")
    hp, popart = _loss_hyper(name)
    L = d["on_reset"].shape[0]
    lo, hi = 0, L - 1
    val, olp, rs = _f32(d, "value", "old_logp", "on_reset")
    adv, ret, nl, vp, en = _f32(d, "adv", "ret", "new_logp", "v_pred", "entropy")
    mask = 1 - rs[lo + 1:hi + 1]
    pa = _popart_from(d["popart_mean_std_after"]) if popart else None
    out = M.ppo_loss_ref(nl, olp[lo:hi], vp, val[lo:hi], ret[lo:hi], adv[lo:hi], en, mask, hp, popart=pa)
    assert_close_ref(out["loss"], d["loss"], tol=1e-6, what="loss")
    msum = float(mask.sum())
    for k in ("g_logp", "g_value", "g_entropy"):
        assert_close_ref(out[k].numpy() * msum, d[k] * msum, tol=1e-6, what=k)
    for k, v in out["stats"].items():
        assert_close_ref(v, d[f"stat_{k}"], tol=1e-6, what=f"stat {k}")
    if popart:  # the PopArt update that ran between GAE and the loss (mappo.py:263-264)
        rms = M.RunningMeanStdRef((1,), beta=float(d["popart_beta"]))
        rms.mean, rms.mean_sq, rms.debias = (torch.tensor([x], dtype=torch.float64) for x in d["popart_state_before"])
        rms.update(ret[lo:hi], mask=mask)
        got = np.array([rms.mean.item(), rms.mean_sq.item(), rms.debias.item()])
        np.testing.assert_allclose(got, d["popart_state_after"], rtol=1e-14)


def test_masknorm_popart_fixture():
    d = load_golden("masknorm_popart.npz")
    adv, mask = torch.from_numpy(d["adv"]), torch.from_numpy(d["mask"])
    assert_close_ref(M.masked_normalization_ref(adv, mask), d["norm_adv"], tol=1e-7)
    assert_close_ref(M.masked_normalization_ref(adv, None), d["norm_adv_nomask"], tol=1e-7)
    rms = M.RunningMeanStdRef((1,), beta=float(d["pa_beta"]))
    y = torch.from_numpy(d["pa_y"])
    for i in range(4):
        rms.update(torch.from_numpy(d["pa_x"][i]), mask=torch.from_numpy(d["pa_mask"][i]) if i % 2 else None)
        got = np.array([rms.mean.item(), rms.mean_sq.item(), rms.debias.item()])
        np.testing.assert_allclose(got, d["pa_states"][i], rtol=1e-14)
        assert_close_ref(rms.normalize(y), d["pa_norm"][i], tol=1e-7)
        assert_close_ref(rms.denormalize(y), d["pa_denorm"][i], tol=1e-7)


def test_stack_fixture():
    """recursive_aggregate(np.stack(axis=1)) incl. zero-fill of leaves missing in some samples."""
    d = load_golden("stack.npz")
    B = 6
    keys = sorted({k.split(".", 1)[1] for k in d if k.startswith("s0.") or k.startswith("s1.")})
    per_sample = [{k: d.get(f"s{b}.{k}") for k in keys} for b in range(B)]
    got = M.stack_leaves_ref(per_sample)
    out_keys = [str(k) for k in d["out_keys"]]
    assert sorted(out_keys) == sorted(k for k, v in got.items() if v is not None)
    for k in out_keys:
        assert got[k].dtype == d[f"out.{k}"].dtype
        assert np.array_equal(got[k], d[f"out.{k}"]), k
    assert (got["truncated"][:, 0::2] == 0).all()  # zero-filled columns (namedarray.py:588-595)


def test_logp_entropy_from_logits_matches_torch_distribution():
    cfg = synth.CONFIGS["cfg5_hns_scale"]
    logits, actions = synth.make_logits_actions(cfg, (6, 5), seed=0)
    lp, en = M.logp_entropy_from_logits_ref(torch.from_numpy(logits), torch.from_numpy(actions).long(), cfg.num_actions)
    z = torch.from_numpy(logits).double()
    off, lp2, en2 = 0, 0, 0
    for h, k in enumerate(cfg.num_actions):
        ls = torch.log_softmax(z[..., off:off + k], -1)
        lp2 = lp2 + ls.gather(-1, torch.from_numpy(actions[..., h:h + 1]).long())
        en2 = en2 - (ls.exp() * ls).sum(-1, keepdim=True)
        off += k
    assert_close_ref(lp, lp2, what="logp")
    assert_close_ref(en, en2, what="entropy")


def test_synth_satisfies_reference_invariants():
    """The data invariants gae.py:69-77 asserts (production disables them with -O)."""
    for name in ("cfg1_atari_cpu", "cfg3_smac_27m"):
        cfg = synth.CONFIGS[name]
        s = synth.make_sample_scalars(cfg, seed=0, B=8)
        tr, dn, rs, rew = (s[k].astype(np.float64) for k in ("truncated", "done", "on_reset", "reward"))
        assert (tr * dn == 0).all()
        assert ((tr + dn)[:-1] == rs[1:]).all()
        assert (rew[:-1] * rs[1:] == 0).all()
        assert s["reward"].dtype == np.float32 and s["done"].dtype == np.uint8


def test_hot_path_ref_runs_minibatched():
    cfg = synth.PathConfig("mini", T=8, B=16, epochs=2, minibatches=4, p_end=0.1)
    s = synth.make_sample_scalars(cfg, 0)
    pol = synth.make_policy_outputs(cfg, s, 1)
    batch = {k: torch.from_numpy(v).float() for k, v in s.items()}
    batch.update({k: torch.from_numpy(v) for k, v in pol.items()})
    out = M.hot_path_ref(batch, M.LossHyper(), cfg.gamma, cfg.lmbda, cfg.epochs, cfg.minibatches, seed=5)
    assert len(out["per_minibatch"]) == 8
    assert out["adv"].shape == (cfg.L, 16, 1) and float(out["adv"][-1].abs().sum()) == 0.0
    assert all(np.isfinite(float(r["loss"])) for r in out["per_minibatch"])


# ---- 3. live differential against the unmodified reference (build container only) ---------------------
needs_reference = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present on this box")


@needs_reference
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_live_gae_and_loss_vs_reference(seed):
    R = ref_loader.load()
    cfg = synth.PathConfig("live", T=20 + seed, B=7 + seed, A=1 + seed, p_end=0.07, gamma=0.98, lmbda=0.93)
    s = synth.make_sample_scalars(cfg, seed)
    pol = synth.make_policy_outputs(cfg, s, seed, epochs=1)
    t = {k: torch.from_numpy(v).float() for k, v in s.items()}
    kw = dict(eps_clip=0.2, clip_value=bool(seed % 2), dual_clip=bool((seed + 1) % 2), value_loss=M.VALUE_LOSS_KINDS[seed],
              value_loss_weight=0.9)
    tr = R.mappo.MultiAgentPPO(ref_loader.FakePolicy(), discount_rate=cfg.gamma, gae_lambda=cfg.lmbda, **kw)
    NA = R.namedarray.NamedArray
    mk = lambda a, r: R.trainer.SampleBatch(obs=None, on_reset=t["on_reset"], done=t["done"], truncated=t["truncated"],
                                            reward=t["reward"],
                                            analyzed_result=NA(value=t["value"], log_probs=t["old_logp"], adv=a, ret=r))
    adv, ret = tr._compute_adv_and_value_target(mk(None, None), None)
    a2, r2 = M.adv_and_value_target_ref(t["reward"], t["value"], t["truncated"], t["done"], t["on_reset"], cfg.gamma,
                                        cfg.lmbda)
    assert torch.equal(adv, a2) and torch.equal(ret, r2)
    adv, ret = M.pad_last_row(adv), M.pad_last_row(ret)
    hi = cfg.L - 1
    mask = 1 - t["on_reset"][1:hi + 1]
    nl, vp, en = (torch.from_numpy(pol[k][0]).requires_grad_(True) for k in ("new_logp", "v_pred", "entropy"))
    loss, _ = tr._compute_loss(mk(adv, ret)[0:hi], R.mappo.SampleAnalyzedResult(t["old_logp"][:hi], nl, vp, en), mask)
    loss.backward()
    mine = M.ppo_loss_ref(nl.detach(), t["old_logp"][:hi], vp.detach(), t["value"][:hi], ret[:hi], adv[:hi], en.detach(),
                          mask, M.LossHyper(**kw))
    assert torch.equal(mine["loss"], loss.detach())
    for k, g in (("g_logp", nl.grad), ("g_value", vp.grad), ("g_entropy", en.grad)):
        assert torch.equal(mine[k], g), k


def test_n_step_return_known_answers_and_fixture():
    """legacy/tests/modules_test.py:180-209 (hand-computed) and the fixture written by the unmodified reference."""
    rew = [1, 2, 1, -3, 0, 3, -1, 2, 1, -2, 10, 1]
    value = [2, 10, 1, 5, 0, 3, 2, -2, -10, 5, -1, -100]
    done = [0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0]
    truncated = [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0]
    args = (torch.tensor(rew[:-1], dtype=torch.float32), torch.tensor(value[1:], dtype=torch.float32),
            torch.tensor(done[1:], dtype=torch.float32), torch.tensor(truncated[1:], dtype=torch.float32))
    ret = M.n_step_return_ref(2, *args, 0.1)
    np.testing.assert_almost_equal(ret[:4].numpy(), [1.21, 2.15, 0.7, -3], decimal=6)
    np.testing.assert_almost_equal(ret[5:9].numpy(), [2.88, -0.9, 2.15, 1.5], decimal=6)
    ret = M.n_step_return_ref(3, *args, 0.1)
    np.testing.assert_almost_equal(ret[:4].numpy(), [1.215, 2.07, 0.7, -3], decimal=6)
    np.testing.assert_almost_equal(ret[5:].numpy(), [2.91, -0.79 + 5e-3, 2.15, 1.5], decimal=6)
    d = load_golden("nstep.npz")
    for name in sorted({k.split(".")[0] for k in d}):
        got = M.n_step_return_ref(int(d[f"{name}.n"]), torch.from_numpy(d[f"{name}.reward"]),
                                  torch.from_numpy(d[f"{name}.nex_value"]), torch.from_numpy(d[f"{name}.nex_done"]).float(),
                                  torch.from_numpy(d[f"{name}.nex_truncated"]).float(), float(d[f"{name}.gamma"]))
        assert np.array_equal(got.numpy(), d[f"{name}.ret"]), name


# ---- GAE family: TrajGAE and the general gae_trace surface ------------------------------------------------
TRAJ_FIXTURES = ["f32w1", "f32w3", "f64w2"]
GENERAL_FIXTURES = ["nc3", "gam", "lam", "both_nc2", "all", "vt_nc1"]


def split_episodes(fx, name):
    lens = fx[f"{name}.lens"]
    lo = np.concatenate([[0], np.cumsum(lens)])
    alo = np.concatenate([[0], np.cumsum(np.maximum(lens - 1, 0))])
    for k, n in enumerate(lens):
        yield (k, fx[f"{name}.reward"][lo[k]:lo[k + 1]], fx[f"{name}.value"][lo[k]:lo[k + 1]],
               fx[f"{name}.adv"][alo[k]:alo[k + 1]], fx[f"{name}.ret"][alo[k]:alo[k + 1]])


@pytest.mark.parametrize("name", TRAJ_FIXTURES)
def test_traj_gae_restatement_matches_reference_fixture(name):
    """Bit-exact against what the unmodified TrajGAE.process wrote (gae.py:100-139), in the arrays' own dtype."""
    fx = load_golden("traj_gae.npz")
    for k, reward, value, adv, ret in split_episodes(fx, name):
        a, r = M.traj_gae_process_ref(reward, value, fx[f"{name}.final_truncated"][k], bool(fx[f"{name}.final_has_value"][k]),
                                      float(fx[f"{name}.gamma"]), float(fx[f"{name}.lmbda"]))
        assert a.dtype == adv.dtype
        assert np.array_equal(a, adv) and np.array_equal(r, ret), f"episode {k}"


def test_traj_gae_restatement_known_answer_integer_arrays():
    """legacy/tests/modules_test.py:140-178 with the test's own int64 arrays (numpy promotes to float64)."""
    a, _ = M.traj_gae_process_ref(np.array([[1], [2], [0]]), np.array([[2], [0], [1]]), np.array([1]), True, 0.1, 0.1)
    np.testing.assert_allclose(a[:, 0], [2.1 * 0.01 - 1, 2.1])
    a, _ = M.traj_gae_process_ref(np.array([[1], [3], [0]]), np.array([[2], [2], [0]]), np.array([0]), True, 0.1, 0.1)
    np.testing.assert_allclose(a[:, 0], [-0.8 + 0.01, 1])


@pytest.mark.parametrize("name", GENERAL_FIXTURES)
def test_gae_trace_restatement_vector_critic_and_tensor_discounts(name):
    """gae_trace_ref against the unmodified gae_trace on vector critics / per-element gamma, lambda / importance ratio."""
    fx = load_golden("gae_general.npz")
    g = lambda k: fx[f"{name}.{k}"]
    as_arg = lambda x: float(x) if x.ndim == 0 else torch.from_numpy(x)
    vt = f"{name}.imp_ratio" in fx
    got = M.gae_trace_ref(torch.from_numpy(g("reward")), torch.from_numpy(g("value")), torch.from_numpy(g("truncated")).float(),
                          torch.from_numpy(g("done")).float(), torch.from_numpy(g("on_reset")).float(), as_arg(g("gamma")),
                          as_arg(g("lmbda")), vtrace=vt, imp_ratio=torch.from_numpy(g("imp_ratio")) if vt else None, rho=1.0,
                          c=0.9)
    assert np.array_equal(got.numpy(), g("adv"))


@pytest.mark.skipif(not ref_loader.available(), reason="live reference only in the build container")
def test_traj_gae_restatement_vs_live_reference():
    R = ref_loader.load()
    rng = np.random.default_rng(3)
    for trial in range(6):
        n, W = int(rng.integers(2, 30)), int(rng.integers(1, 4))
        reward = rng.standard_normal((n, W)).astype(np.float32)
        value = rng.standard_normal((n, W)).astype(np.float32)
        trunc = np.full((W,), trial % 2, np.uint8)
        has = trial % 3 != 0
        memory = [R.trainer.SampleBatch(obs=None, reward=reward[i].copy(), done=1 - trunc, truncated=trunc,
                                        analyzed_result=None if (i == n - 1 and not has) else
                                        R.namedarray.NamedArray(value=value[i].copy(), adv=None, ret=None))
                  for i in range(n)]
        memory = R.gae.TrajGAE(0.98, 0.9).process(memory)
        a, r = M.traj_gae_process_ref(reward, value, trunc, has, 0.98, 0.9)
        assert np.array_equal(a, np.stack([m.analyzed_result.adv for m in memory[:-1]]))
        assert np.array_equal(r, np.stack([m.analyzed_result.ret for m in memory[:-1]]))


def test_traj_gae_and_gae_trace_restatements_agree_on_one_episode():
    """Two reference functions, one recurrence: for a single finished episode in float64, TrajGAE (gae.py:100-139) and
    gae_trace (gae.py:8-97, float64 inside) compute the same products and sums (a + b == b + a), so their restatements
    must agree to the bit once gae_trace's result is compared before / after the same float32 cast."""
    rng = np.random.default_rng(8)
    for truncated_end in (False, True):
        n = 23
        reward = rng.standard_normal((n, 1))
        value = rng.standard_normal((n, 1))
        reward[-1] = 0.0  # the final step carries no reward (gae.py:72)
        a, _ = M.traj_gae_process_ref(reward, value, np.array([float(truncated_end)]), True, 0.97, 0.9)
        # the same episode as a [n+1]-row batch lane: the row after the final step is a reset row
        done = np.zeros((n + 1, 1)); trunc = np.zeros((n + 1, 1)); reset = np.zeros((n + 1, 1))
        (trunc if truncated_end else done)[n - 1] = 1.0
        reset[n] = 1.0
        # TrajGAE bootstraps the last computed step with value[last] * truncated[last]; gae_trace reads value[t+1] as is,
        # and mappo zeroes it where done (mappo.py:120-124)
        v = np.concatenate([value, np.zeros((1, 1))]) * (1 - done)
        got = M.gae_trace_ref(*(torch.from_numpy(x) for x in (reward, v, trunc, done, reset)), 0.97, 0.9)
        # rows 0 .. n-2 are TrajGAE's steps; row n-1 is the final step itself (delta = -v there if truncated)
        assert np.array_equal(got.numpy()[:n - 1], a.astype(np.float32))
