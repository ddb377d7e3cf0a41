"""GPU parity tests of the rest of the GAE family and the chunk reshapes, all through the C-ABI:
  srl_gae_trace   vs the unmodified modules.gae_trace (tests/golden/gae_general.npz) and vs srl_gae_scan -- bit-exact
  srl_traj_gae    vs the unmodified TrajGAE.process (tests/golden/traj_gae.npz) and the reference's own KAT -- bit-exact
  to_chunk / back_to_trajectory (K1 launches) vs torch.cat(torch.split) and the reference's KAT -- bit-exact
"""
import numpy as np
import pytest
import torch

from oracle import ref_math as M
from srl_b200 import synth
from tests.test_oracle import GENERAL_FIXTURES, TRAJ_FIXTURES, split_episodes
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback to test instead")
    from srl_b200 import ops as _ops
    _ops._lib.load_library()
    return _ops


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# ---- srl_gae_trace ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GENERAL_FIXTURES)
def test_gae_trace_matches_reference_fixture(ops, name):
    fx = load_golden("gae_general.npz")
    g = lambda k: fx[f"{name}.{k}"]
    as_arg = lambda x: float(x) if x.ndim == 0 else dev(x)
    vt = f"{name}.imp_ratio" in fx
    adv = ops.gae_trace(dev(g("reward")), dev(g("value")), dev(g("truncated")), dev(g("done")), dev(g("on_reset")),
                        as_arg(g("gamma")), as_arg(g("lmbda")), vtrace=vt, imp_ratio=dev(g("imp_ratio")) if vt else None,
                        rho=1.0, c=0.9)
    torch.cuda.synchronize()
    assert adv.shape == g("adv").shape
    assert np.array_equal(adv.cpu().numpy(), g("adv"))


@pytest.mark.parametrize("cfg_name,B", [("cfg1_atari_cpu", None), ("cfg3_smac_27m", 24), ("cfg2_atari_large", 320)])
def test_gae_trace_equals_gae_scan(ops, cfg_name, B):
    """With a scalar critic and scalar discounts both entry points run the same arithmetic: identical bits, including
    the zero padding row and the value target."""
    import dataclasses
    cfg = synth.CONFIGS[cfg_name]
    if B is not None:
        cfg = dataclasses.replace(cfg, B=B)
    s = synth.make_sample_scalars(cfg, 4)
    d = {k: dev(v) for k, v in s.items()}
    adv1, ret1, _ = ops.gae_scan(d["reward"], d["value"], d["done"], d["truncated"], d["on_reset"], cfg.gamma, cfg.lmbda)
    adv2, ret2 = ops.gae_trace(d["reward"], d["value"], d["truncated"], d["done"], d["on_reset"], cfg.gamma, cfg.lmbda,
                               apply_done=True, want_ret=True, pad_last_row=True)
    torch.cuda.synchronize()
    assert torch.equal(adv1, adv2) and torch.equal(ret1, ret2)


def test_gae_trace_vector_critic_is_per_column_scalar_scan(ops):
    """critic_dim = 5 at a size that spans many CTAs: column k of the vector result == the scalar scan of column k
    (the flags broadcast over the critic axis, gae.py:26-30); the oracle checks one column end to end."""
    T, B, Nc = 47, 333, 5
    cfg = synth.PathConfig("vc", T=T, B=B, p_end=0.05)
    s = synth.make_sample_scalars(cfg, 9)
    g = torch.Generator().manual_seed(1)
    reward = torch.randn(T + 1, B, Nc, generator=g)
    value = torch.randn(T + 1, B, Nc, generator=g)
    fl = {k: dev(s[k]) for k in ("done", "truncated", "on_reset")}
    adv, ret = ops.gae_trace(reward.cuda(), value.cuda(), fl["truncated"], fl["done"], fl["on_reset"], 0.99, 0.95,
                             apply_done=True, want_ret=True)
    for k in (0, 3, 4):
        a1, r1 = ops.gae_trace(reward[..., k:k + 1].contiguous().cuda(), value[..., k:k + 1].contiguous().cuda(),
                               fl["truncated"], fl["done"], fl["on_reset"], 0.99, 0.95, apply_done=True, want_ret=True)
        assert torch.equal(adv[..., k:k + 1], a1) and torch.equal(ret[..., k:k + 1], r1)
    f32 = {k: torch.from_numpy(s[k]).float() for k in ("done", "truncated", "on_reset")}
    ra, rr = M.adv_and_value_target_ref(reward[..., 2:3], value[..., 2:3], f32["truncated"], f32["done"], f32["on_reset"],
                                        0.99, 0.95)
    assert torch.equal(adv[..., 2:3].cpu(), ra) and torch.equal(ret[..., 2:3].cpu(), rr)


def test_gae_trace_rejects_bad_arguments(ops):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device="cuda")
    u = lambda *s: z(*s, dt=torch.uint8)
    with pytest.raises(AssertionError):  # gae.py:52: gamma must be a float or a FloatTensor
        ops.gae_trace(z(4, 3, 1), z(5, 3, 1), u(5, 3, 1), u(5, 3, 1), u(5, 3, 1), 1, 0.9)
    with pytest.raises(AssertionError):  # gae.py:53: gamma tensor shaped like on_reset[:-1]
        ops.gae_trace(z(4, 3, 1), z(5, 3, 1), u(5, 3, 1), u(5, 3, 1), u(5, 3, 1), z(5, 3, 1), 0.9)
    with pytest.raises(ValueError):
        ops.gae_trace(z(4, 3, 1).cpu(), z(5, 3, 1), u(5, 3, 1), u(5, 3, 1), u(5, 3, 1), 0.9, 0.9)
    with pytest.raises(ValueError):
        ops.gae_trace(z(4, 3, 1), z(5, 3, 1), u(5, 3, 1), u(5, 3, 1), u(5, 3, 1), 0.9, 0.9, vtrace=True)


# ---- srl_traj_gae ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", TRAJ_FIXTURES)
def test_traj_gae_matches_reference_fixture(ops, name):
    """All episodes of a fixture in ONE launch; bit-exact in the arrays' own dtype."""
    fx = load_golden("traj_gae.npz")
    lens = fx[f"{name}.lens"]
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    adv, ret = ops.traj_gae(dev(fx[f"{name}.reward"]), dev(fx[f"{name}.value"]), dev(offsets),
                            dev(fx[f"{name}.final_truncated"]), dev(fx[f"{name}.final_has_value"]),
                            float(fx[f"{name}.gamma"]), float(fx[f"{name}.lmbda"]))
    torch.cuda.synchronize()
    adv, ret = adv.cpu().numpy(), ret.cpu().numpy()
    assert adv.dtype == fx[f"{name}.adv"].dtype
    for k, _, _, ref_adv, ref_ret in split_episodes(fx, name):
        lo, n = offsets[k], lens[k]
        assert np.array_equal(adv[lo:lo + n - 1], ref_adv) and np.array_equal(ret[lo:lo + n - 1], ref_ret), f"episode {k}"
        assert not adv[lo + n - 1].any() and not ret[lo + n - 1].any()  # final step untouched


def _memory(rew, value, done, truncated, final_value=True):
    from srl_b200.api import AnalyzedResult, SampleBatch
    n = len(rew)
    return [SampleBatch(obs=None, reward=np.array([r]),
                        analyzed_result=AnalyzedResult(value=np.array([v])) if (final_value or i < n - 1) else None,
                        done=np.array([d]), truncated=np.array([t]))
            for i, (r, v, d, t) in enumerate(zip(rew, value, done, truncated))]


def test_traj_gae_postprocessor_reference_known_answers(ops):
    """legacy/tests/modules_test.py:140-178 through the registered post-processor (int64 arrays -> float64)."""
    from srl_b200 import api, postprocess  # noqa: F401  (registers 'gae_b200')
    proc = api.make_traj_postprocessor(type("Cfg", (), dict(type_="gae_b200", args=dict(gamma=0.1, lmbda=0.1)))())
    memory = proc.process(_memory([1, 2, 0], [2, 0, 1], [0, 0, 0], [0, 0, 1]))
    for a, b in zip([m.analyzed_result.adv.item() for m in memory[:-1]], [2.1 * 0.01 - 1, 2.1]):
        assert a == pytest.approx(b, abs=1e-7)
    memory = proc.process(_memory([1, 3, 0], [2, 2, 0], [0, 0, 1], [0, 0, 0]))
    for a, b in zip([m.analyzed_result.adv.item() for m in memory[:-1]], [-0.8 + 0.01, 1]):
        assert a == pytest.approx(b, abs=1e-7)
    assert memory[-1].analyzed_result.adv is None  # the final step is not written (gae.py:113)
    with pytest.raises(AssertionError):  # gae.py:110: the last step must end the episode
        proc.process(_memory([1, 2], [1, 1], [0, 0], [0, 0]))


def test_traj_gae_postprocessor_many_episodes_vs_oracle(ops):
    """process_many: 200 float32 episodes of ragged lengths (width 2) in one launch vs the numpy restatement."""
    from srl_b200.api import AnalyzedResult, SampleBatch
    from srl_b200.postprocess import TrajGAEB200
    rng = np.random.default_rng(0)
    memories, refs = [], []
    for k in range(200):
        n = int(rng.integers(1, 60))
        reward = rng.standard_normal((n, 2, 1)).astype(np.float32)
        value = rng.standard_normal((n, 2, 1)).astype(np.float32)
        trunc = np.full((2, 1), k % 2, np.uint8)
        has = k % 5 != 0
        memories.append([SampleBatch(obs=None, reward=reward[i], done=1 - trunc, truncated=trunc,
                                     analyzed_result=None if (i == n - 1 and not has) else AnalyzedResult(value=value[i]))
                         for i in range(n)])
        refs.append(M.traj_gae_process_ref(reward, value, trunc, has, 0.99, 0.95))
    out = TrajGAEB200(0.99, 0.95).process_many(memories)
    for memory, (ra, rr) in zip(out, refs):
        for i in range(len(memory) - 1):
            ar = memory[i].analyzed_result
            assert ar.adv.dtype == np.float32 and ar.adv.shape == (2, 1)
            assert np.array_equal(ar.adv, ra[i]) and np.array_equal(ar.ret, rr[i])


# ---- to_chunk / back_to_trajectory -------------------------------------------------------------------------------
def test_to_chunk_reference_known_answer(ops):
    """legacy/tests/modules_test.py:79-89."""
    x = torch.randn(24, 8, 1)
    x_ = ops.to_chunk(x.cuda(), 8)
    assert torch.equal(x_.cpu(), x.view(8, 3, 8, 1).transpose(0, 1).reshape(3, 64, 1))
    assert torch.equal(ops.back_to_trajectory(x_, 8).cpu(), x)


@pytest.mark.parametrize("shape,C,dtype", [((24, 8, 1), 8, torch.float32), ((128, 96, 1), 16, torch.uint8),
                                             ((40, 33, 7), 5, torch.float32), ((16, 64, 4, 84, 84), 4, torch.uint8),
                                             ((6, 3), 1, torch.float64), ((12, 5, 3), 12, torch.int32)])
def test_to_chunk_bit_exact(ops, shape, C, dtype):
    g = torch.Generator().manual_seed(2)
    x = torch.randint(0, 255, shape, generator=g).to(dtype)
    ref = torch.cat(torch.split(x, shape[0] // C, dim=0), dim=1)  # utils.py:180
    got = ops.to_chunk(x.cuda(), C)
    assert got.shape == ref.shape and torch.equal(got.cpu(), ref)
    back = ops.back_to_trajectory(got, C)
    assert torch.equal(back.cpu(), torch.cat(torch.split(ref, ref.shape[1] // C, dim=1), dim=0))  # utils.py:195
    assert torch.equal(back.cpu(), x)


def test_to_chunk_rejects_indivisible_time_axis(ops):
    with pytest.raises(IndexError):  # utils.py:176-179
        ops.to_chunk(torch.zeros(10, 4, 1, device="cuda"), 3)
    with pytest.raises(ValueError):
        ops.to_chunk(torch.zeros(10, 4, 1), 2)


@pytest.mark.parametrize("T,B,C,n,D", [(12, 8, 3, 4, (5,)), (16, 32, 4, 8, (4, 6, 6)), (8, 16, 8, 16, ()), (10, 6, 1, 3, (2,))])
def test_gather_to_chunk_equals_gather_then_to_chunk(T, B, C, n, D):
    """(f)4: the minibatch gather and the RNN-chunk reshape in ONE K1 launch == modules.to_chunk(x[:, env_idx], C)
    (legacy/algorithm/modules/utils.py:164-180), bit for bit, for float32 and uint8 leaves."""
    from srl_b200 import ops
    g = torch.Generator().manual_seed(T * 100 + B)
    idx = torch.randperm(B, generator=g)[:n].to(torch.int32)
    for dtype in (torch.float32, torch.uint8):
        x = (torch.rand((T, B) + D, generator=g) * 200).to(dtype)
        want = torch.cat(torch.split(x[:, idx.long()], T // C, dim=0), dim=1)  # the reference's to_chunk on the gathered leaf
        got = ops.gather_to_chunk(x.cuda(), idx.cuda(), C)
        assert got.dtype == dtype and tuple(got.shape) == tuple(want.shape)
        assert torch.equal(got.cpu(), want)
    with pytest.raises(IndexError, match="must be a multiple of"):
        ops.gather_to_chunk(torch.zeros(7, 4, 2).cuda(), idx[:2].cuda() % 4, 2)


@pytest.mark.parametrize("N,n_env,group,epochs", [(4096, 4096, 1, 4), (1024, 100, 3, 2), (512, 37, 1, 5), (64, 64, 1, 2), (20480, 2048, 10, 1)])
def test_scan_with_permutation_on_the_side_equals_philox_perm(N, n_env, group, epochs):
    """srl_gae_scan_perm: the permutations computed by the scan kernel's worker threads (warp-specialised kernel: 256 <= N <
    ~9500 lanes here) or by the stand-alone kernel behind the scan (the other shapes) are srl_philox_perm's, bit for bit, and
    the scan's own outputs do not change."""
    from srl_b200 import ops
    L = 33
    g = torch.Generator().manual_seed(N + n_env)
    f = lambda: torch.randn((L, N), generator=g).cuda()
    u = lambda p: (torch.rand((L, N), generator=g) < p).to(torch.uint8).cuda()
    reward, value, done, trunc, reset = f(), f(), u(0.02), u(0.01), u(0.03)
    want = ops.philox_perm(77, 3, n_env, group, n_epochs=epochs)
    a0, r0, p0 = ops.gae_scan(reward, value, done, trunc, reset, 0.99, 0.95)
    out = torch.full((epochs, n_env * group), -7, dtype=torch.int32, device="cuda")
    job = dict(seed=77, epoch=3, n_epochs=epochs, n_env=n_env, group=group, out=out)
    a1, r1, p1 = ops.gae_scan(reward, value, done, trunc, reset, 0.99, 0.95, perm_job=job)
    torch.cuda.synchronize()
    assert torch.equal(out, want.view(epochs, -1))
    assert torch.equal(a0, a1) and torch.equal(r0, r1) and torch.equal(p0, p1)
    assert job["fused"] == (256 <= N < 2 * 148 * 32)


@pytest.mark.parametrize("N,n_env,group,epochs,mbs", [(4096, 4096, 1, 4, 8), (1024, 128, 8, 2, 4), (272, 34, 8, 8, 2),
                                                      (1536, 512, 3, 3, 1), (512, 512, 1, 9, 2), (20480, 2048, 10, 1, 4)])
def test_scan_minibatch_shares_add_up_to_the_minibatch_sums(N, n_env, group, epochs, mbs):
    """srl_gae_scan_perm's `minibatch_part`: table[slot][cta] is CTA cta's 32 lanes' share of minibatch slot's {count, sum,
    sum of squares} -- exactly the lanes of that 32-lane group which the permutation of the SAME launch puts into the
    minibatch.  Not written (part_valid False, table untouched) by the kernels that do not compute the permutation or for
    more than 8 epochs."""
    from srl_b200 import ops
    L = 21
    g = torch.Generator().manual_seed(N + n_env + mbs)
    f = lambda: torch.randn((L, N), generator=g).cuda()
    u = lambda p: (torch.rand((L, N), generator=g) < p).to(torch.uint8).cuda()
    reward, value, done, trunc, reset = f(), f(), u(0.02), u(0.01), u(0.05)
    perm = torch.empty((epochs, N), dtype=torch.int32, device="cuda")
    part = ops.new_minibatch_part(N, "cuda").fill_(-3.0)
    aos = torch.empty((N, 4), dtype=torch.float64, device="cuda")
    pack = ops.new_pack(L, N, "cuda")
    job = dict(seed=5, epoch=1, n_epochs=epochs, n_env=n_env, group=group, out=perm, minibatches=mbs, part=part)
    _, _, lane_part = ops.gae_scan(reward, value, done, trunc, reset, 0.99, 0.95, old_logp=f(), pack=pack, lane_aos=aos,
                                   perm_job=job)
    torch.cuda.synchronize()
    assert torch.equal(perm, ops.philox_perm(5, 1, n_env, group, n_epochs=epochs).view(epochs, -1))
    want_valid = (256 <= N < 2 * 148 * 32) and epochs <= 8 and epochs * mbs <= 32
    assert job["part_valid"] == want_valid
    if not want_valid:
        assert bool((part == -3.0).all())
        return
    lp, pm, tb = lane_part.cpu().numpy(), perm.cpu().numpy(), part.cpu().numpy()
    ctas, per = (N + 31) // 32, N // mbs
    assert bool((tb[epochs * mbs:] == -3.0).all()) and not tb[:epochs * mbs, :, 3].any()
    for e in range(epochs):
        for j in range(mbs):
            lanes = pm[e, j * per:(j + 1) * per]
            want = np.zeros((ctas, 3))
            for k in range(3):
                np.add.at(want[:, k], lanes // 32, lp[k, lanes])
            np.testing.assert_allclose(tb[e * mbs + j, :, :3], want, rtol=1e-14, atol=1e-14)
            assert np.array_equal(tb[e * mbs + j, :, 0], np.bincount(lanes // 32, weights=lp[0, lanes], minlength=ctas))


@pytest.mark.parametrize("T,B,C,n,layers,H", [(8, 6, 2, 4, 1, 16), (12, 40, 4, 17, 2, 33), (400, 64, 40, 32, 1, 64),
                                              (10, 9, 1, 9, 1, 8), (6, 300, 3, 300, 1, 4)])
def test_rnn_chunk_prep_equals_the_reference_expressions(ops, T, B, C, n, layers, H):
    """(f)4: reset flags chunked, AutoResetRNN's segment boundaries and the masked, layer-major chunk-start hidden states in
    one launch == to_chunk(x[:, env_idx], C) (utils.py:164-180), `(masks[1:] == 0).any(dim=1).nonzero()`
    (autoreset_rnn.py:46-47) and `x[0].transpose(0, 1)` times `masks[0]` (actor_critic_policy.py:362-363,
    autoreset_rnn.py:55), bit for bit (signed zeros included)."""
    g = torch.Generator().manual_seed(T * 7 + B)
    to_chunk = lambda x: torch.cat(torch.split(x, T // C, dim=0), dim=1)
    idx = torch.randperm(B, generator=g)[:n].to(torch.int32)
    on_reset = (torch.rand((T, B, 1), generator=g) < (0.5 / max(1, n * C // 4))).to(torch.uint8)
    on_reset[0, idx[0].item()] = 1  # a reset in the first row: its chunk-start hidden state is masked
    hx = torch.randn((T, B, layers, H), generator=g)
    sample_reset = to_chunk(on_reset[:, idx.long()].float())  # [Tc, C*n, 1] as the policy sees it
    masks = 1 - sample_reset
    has_zeros = (masks[1:] == 0.0).any(dim=1).nonzero(as_tuple=True)[0].numpy()
    want_segments = [0] + (has_zeros + 1).tolist() + [T // C]
    want_hx0 = to_chunk(hx[:, idx.long()])[0].transpose(0, 1) * masks[0].view(1, -1, 1)
    got_reset, row_any, got_hx0 = ops.rnn_chunk_prep(on_reset.cuda(), idx.cuda(), C, hx=hx.cuda())
    assert torch.equal(got_reset.cpu(), sample_reset[..., 0].to(torch.uint8))
    assert ops.reset_segments(row_any) == want_segments
    assert int(row_any[0]) == int(sample_reset[0].any())
    assert got_hx0.shape == want_hx0.shape
    assert torch.equal(got_hx0.cpu().view(torch.int32), want_hx0.contiguous().view(torch.int32))
    only_flags = ops.rnn_chunk_prep(on_reset.cuda(), idx.cuda(), C)
    assert only_flags[2] is None and torch.equal(only_flags[0], got_reset) and torch.equal(only_flags[1], row_any)
    with pytest.raises(IndexError, match="must be a multiple of"):
        ops.rnn_chunk_prep(torch.zeros((7, 4), dtype=torch.uint8).cuda(), idx[:2].cuda() % 4, 2)
