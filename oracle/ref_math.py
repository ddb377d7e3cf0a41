"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch-CPU / numpy) of SRL's trainer hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this file.  Nothing under ``srl_b200/`` imports it: the product path has no CPU
fallback and raises when the CUDA library is missing.

Parity status
-------------
* GAE, masked normalisation, PopArt / RunningMeanStd, stacking: PINNED.  Checked against the
  reference's own known-answer tests (legacy/tests/modules_test.py:91-138,267-271,301-323) and
  against outputs of the unmodified reference loaded with ``oracle/ref_loader.py`` -- see
  ``tests/test_oracle.py`` and the fixtures written by ``oracle/make_golden.py``.
* PPO loss value + gradients: no reference test pins them.  PINNED here by
  differential fixtures generated from the unmodified ``MultiAgentPPO._compute_loss`` + autograd
  (``tests/golden/loss_*.npz``).
* Minibatch permutation: the feature does not exist in the reference ->
  "parity unpinned" against the reference; the Philox4x32-10 core is pinned to the published
  Random123 known-answer vectors and the gather is pinned to numpy fancy indexing.

Every function cites the reference lines it restates (paths relative to /root/reference).
The op order of the reference is kept on purpose: the CPU baseline in bench.py times these
functions, so they must cost what the reference costs (python loop over T, fp64 casts, ...).
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

# --------------------------------------------------------------------------------------------
# A3: GAE  (legacy/algorithm/modules/gae.py:8-97, legacy/algorithm/ppo/mappo.py:118-144)
# --------------------------------------------------------------------------------------------


@torch.no_grad()
def gae_trace_ref(reward, value, truncated, done, on_reset, gamma, lmbda, vtrace=False,
                  imp_ratio=None, rho=1.0, c=1.0):
    """gae.py:46-97 with high_precision=True: everything promoted to float64, python scan over
    time, result cast to float32.  Shapes: reward [L-1,*], the others [L,*]."""
    f64 = torch.float64
    reward, value, truncated, done, on_reset = (x.to(f64) for x in (reward, value, truncated, done,
                                                                      on_reset))
    if not isinstance(gamma, float):
        gamma = gamma.to(f64)  # gae.py:51-55 per-element discount
    if not isinstance(lmbda, float):
        lmbda = lmbda.to(f64)  # gae.py:56-60
    nxt_alive = 1 - on_reset[1:]
    # gae.py:63 -- note the association: ((gamma * v[1:]) * (1-reset)), then r + that, then - v[:-1]
    delta = reward + gamma * value[1:] * nxt_alive - value[:-1]
    # gae.py:87
    carry = gamma * lmbda * nxt_alive * (1 - truncated[1:])
    if vtrace:
        ratio = imp_ratio.to(f64)
        delta = delta * ratio.clip(max=rho)  # gae.py:64-65
        carry = carry * ratio.clip(max=c)  # gae.py:88-89
    steps = int(reward.shape[0])
    acc = torch.zeros_like(reward[0])
    out = torch.zeros_like(reward)
    for t in range(steps - 1, -1, -1):  # gae.py:91-95
        acc = delta[t] + carry[t] * acc
        out[t] = acc
    return out.float()


@torch.no_grad()
def adv_and_value_target_ref(reward, value, truncated, done, on_reset, gamma, lmbda,
                             popart: Optional["RunningMeanStdRef"] = None, vtrace=False,
                             new_logp=None, old_logp=None):
    """mappo.py:118-144.  All inputs float32 tensors as left by the prefetcher (api/trainer.py:217).
    Returns (adv, ret) of length L-1 (before the zero-row padding of mappo.py:254-256)."""
    base = popart.denormalize(value) if popart is not None else value  # mappo.py:120-124
    boot = base * (1 - done)
    ratio = (new_logp - old_logp).exp() if vtrace else None  # mappo.py:129-132
    adv = gae_trace_ref(reward[:-1], boot, truncated, done, on_reset, gamma, lmbda, vtrace=vtrace,
                        imp_ratio=ratio)
    return adv, adv + boot[:-1]  # mappo.py:143


@torch.no_grad()
def n_step_return_ref(n, reward, nex_value, nex_done, nex_truncated, gamma):
    """legacy/algorithm/modules/n_step_return.py:11-50 (high_precision=True): float64 window sums in the reference's
    operation order; inputs [n+T-1, ...], output [T, ...] float32."""
    reward, nex_value, nex_done, nex_truncated = (x.to(torch.float64) for x in (reward, nex_value, nex_done, nex_truncated))
    T = nex_value.shape[0] - n + 1
    assert T >= 1
    ret = torch.zeros_like(reward[:T])
    discount = torch.ones_like(reward[:T])
    for i in range(n):
        ret += reward[i:i + T] * discount
        # a truncated next step bootstraps its value right away; `discount` is 0 afterwards
        ret += discount * gamma * nex_truncated[i:i + T] * nex_value[i:i + T]
        discount *= gamma * (1 - nex_done[i:i + T]) * (1 - nex_truncated[i:i + T])
    return (ret + discount * nex_value[n - 1:n - 1 + T]).float()


def pad_last_row(x):
    """mappo.py:254-256: F.pad with one zero row at the end of the time axis."""
    return torch.cat([x, torch.zeros_like(x[:1])], dim=0)


def traj_gae_ref(rewards, values, last_truncated: bool, last_has_value: bool, gamma, lmbda):
    """gae.py:100-139 (TrajGAE.process) on plain per-step scalars/arrays.
    rewards/values: sequences of length n (the whole episode incl. the final step)."""
    n = len(rewards)
    adv = [None] * (n - 1)
    ret = [None] * (n - 1)
    acc = np.zeros_like(np.asarray(rewards[0], dtype=np.float64))
    for t in range(n - 2, -1, -1):
        if t == n - 2:
            nxt = (np.asarray(values[t + 1]) * float(last_truncated)) if last_has_value else 0
        else:
            nxt = values[t + 1]
        d = rewards[t] + gamma * nxt - values[t]
        acc = gamma * lmbda * acc + d
        adv[t] = acc
        ret[t] = acc + values[t]
    return adv, ret


def traj_gae_process_ref(reward, value, last_truncated, last_has_value: bool, gamma: float, lmbda: float):
    """gae.py:106-139 on one episode given as arrays: reward / value `[n, W]` in their own dtype (numpy semantics:
    float32 arrays stay float32, python-float gamma and gamma * lmbda are rounded to the array dtype by the ufunc),
    last_truncated `[W]`.  Returns (adv, ret) `[n-1, W]`; the loop and its op order are the reference's."""
    n = reward.shape[0]
    adv, ret = [None] * max(n - 1, 0), [None] * max(n - 1, 0)
    gae = np.zeros_like(reward[0])  # gae.py:112
    step = n - 2
    while step >= 0:
        if step == n - 2:
            bootstrap = value[step + 1] * last_truncated if last_has_value else 0  # gae.py:117-123
        else:
            bootstrap = value[step + 1]
        delta = reward[step] + gamma * bootstrap - value[step]  # gae.py:127
        gae = gamma * lmbda * gae + delta  # gae.py:128
        adv[step] = gae
        ret[step] = gae + value[step]
        step -= 1
    if n < 2:
        empty = np.zeros((0,) + reward.shape[1:], dtype=np.result_type(reward.dtype, value.dtype))
        return empty, empty.copy()
    return np.stack(adv), np.stack(ret)


# --------------------------------------------------------------------------------------------
# A4: masked normalisation (legacy/algorithm/modules/utils.py:10-67)
# --------------------------------------------------------------------------------------------


@torch.no_grad()
def masked_sums_ref(x, mask):
    """The three float64 sums utils.py:54-57 reduces (and all-reduces, :58-61)."""
    xd = x.to(torch.float64)
    if mask is None:
        cnt = torch.tensor(float(xd.numel()), dtype=torch.float64)
    else:
        md = mask.to(torch.float64)
        xd = xd * md
        cnt = md.sum()
    return cnt, xd.sum(), xd.square().sum()


@torch.no_grad()
def masked_normalization_ref(x, mask=None, eps=1e-5, unbiased=False, global_sums=None):
    """utils.py:38-67, dim=None.  ``global_sums=(cnt, s1, s2)`` plays the role of the three
    all-reduced scalars when emulating several ranks on one process (utils.py:58-61)."""
    xd = x.to(torch.float64).clone()
    if mask is not None:
        xd = xd * mask.to(torch.float64)  # masked positions become 0 *before* centring (F5)
    cnt, s1, s2 = masked_sums_ref(x, mask) if global_sums is None else global_sums
    mean = s1 / cnt
    var = s2 / cnt - mean**2  # biased, utils.py:63-64
    if unbiased:
        var = var * (cnt / (cnt - 1))
    return ((xd - mean) / (var.sqrt() + eps)).float()  # eps outside the sqrt, utils.py:67


# --------------------------------------------------------------------------------------------
# A8: RunningMeanStd / PopArt (utils.py:70-151, popart.py:42-59)
# --------------------------------------------------------------------------------------------


class RunningMeanStdRef:
    """Debiased EMA of mean and mean-square in float64 (utils.py:70-151)."""

    def __init__(self, shape=(1,), beta=0.99999, epsilon=1e-5):
        self.beta, self.eps, self.shape = beta, epsilon, tuple(shape)
        self.mean = torch.zeros(self.shape, dtype=torch.float64)
        self.mean_sq = torch.zeros(self.shape, dtype=torch.float64)
        self.debias = torch.zeros(1, dtype=torch.float64)

    @torch.no_grad()
    def batch_sums(self, x, mask=None):
        """utils.py:108-120: (factor, sum x, sum x^2) over all leading dims."""
        xd = x.to(torch.float64)
        lead = tuple(range(xd.dim() - len(self.shape)))
        if mask is None:
            cnt = torch.tensor(float(np.prod(xd.shape[:len(lead)])), dtype=torch.float64)
        else:
            md = mask.to(torch.float64)
            xd = xd * md
            cnt = md.sum()
        return cnt, xd.sum(dim=lead), xd.square().sum(dim=lead)

    @torch.no_grad()
    def update(self, x, mask=None, global_sums=None):
        cnt, s1, s2 = self.batch_sums(x, mask) if global_sums is None else global_sums
        b = self.beta
        self.mean = b * self.mean + (s1 / cnt) * (1.0 - b)  # utils.py:128
        self.mean_sq = b * self.mean_sq + (s2 / cnt) * (1.0 - b)  # utils.py:129
        self.debias = b * self.debias + 1.0 - b  # utils.py:130

    @torch.no_grad()
    def mean_std(self):
        d = self.debias.clamp(min=self.eps)
        m = self.mean / d
        var = (self.mean_sq / d - m**2).clamp(min=1e-2)  # utils.py:136
        return m, var.sqrt()

    @torch.no_grad()
    def normalize(self, x):
        m, s = self.mean_std()
        return ((x.to(torch.float64) - m) / s).clip(-5, 5).float()  # utils.py:139-144

    @torch.no_grad()
    def denormalize(self, x):
        m, s = self.mean_std()
        return (x.to(torch.float64) * s + m).float()  # utils.py:146-151


# --------------------------------------------------------------------------------------------
# A5: PPO / MAPPO loss (mappo.py:146-217, utils.py:228-265)
# --------------------------------------------------------------------------------------------

VALUE_LOSS_KINDS = ("mse", "huber", "smoothl1")


def _pointwise_value_loss(kind, cfg):
    """utils.py:241-265: torch.nn.{MSELoss,HuberLoss,SmoothL1Loss}(reduction='none', **cfg)."""
    cfg = dict(cfg or {})
    if kind == "mse":
        return torch.nn.MSELoss(reduction="none", **cfg)
    if kind == "huber":
        return torch.nn.HuberLoss(reduction="none", **cfg)
    if kind == "smoothl1":
        return torch.nn.SmoothL1Loss(reduction="none", **cfg)
    raise ValueError(kind)


@dataclasses.dataclass
class LossHyper:
    """Hyper-parameters read by mappo.py:71-112 that enter _compute_loss."""
    eps_clip: float = 0.2
    clip_value: bool = False
    value_eps_clip: Optional[float] = None  # defaults to eps_clip (mappo.py:90)
    dual_clip: bool = True
    c_clip: float = 3.0
    value_loss: str = "mse"
    value_loss_config: Optional[dict] = None
    value_loss_weight: float = 0.5
    entropy_bonus_weight: float = 0.01
    normalize_old_value: bool = False

    def veps(self):
        return self.eps_clip if self.value_eps_clip is None else self.value_eps_clip


def ppo_loss_ref(new_logp, old_logp, values, old_value, ret, adv, entropy, mask, hp: LossHyper,
                 popart: Optional[RunningMeanStdRef] = None, global_sums=None, want_grads=True):
    """mappo.py:146-217 followed by ``loss.backward()`` (mappo.py:274).

    All tensors [T, ..., 1] float32 on CPU; ``mask`` = 1 - on_reset[t+1] (mappo.py:260-261).
    Returns dict(loss, policy_loss, value_loss, entropy_loss, g_logp, g_value, g_entropy, stats).
    """
    new_logp = new_logp.detach().clone().requires_grad_(want_grads)
    values = values.detach().clone().requires_grad_(want_grads)
    entropy = entropy.detach().clone().requires_grad_(want_grads)

    prev_v = popart.normalize(old_value) if hp.normalize_old_value else old_value  # mappo.py:151-152
    ratio = (new_logp - old_logp).exp()  # mappo.py:157-158
    denorm_target = None
    target = ret
    if popart is not None:  # mappo.py:174-177
        denorm_target = ret
        target = popart.normalize(ret)

    pointwise = _pointwise_value_loss(hp.value_loss, hp.value_loss_config)
    if hp.clip_value:  # utils.py:228-239
        plain = pointwise(values, target)
        near = prev_v + (values - prev_v).clamp(-hp.veps(), hp.veps())
        v_elem = torch.max(plain, pointwise(near, target))
    else:
        v_elem = pointwise(values, target)
    msum = mask.sum()
    value_loss = (v_elem * mask).sum() / msum  # mappo.py:184

    nadv = masked_normalization_ref(adv, mask, global_sums=global_sums)  # mappo.py:187
    s1 = ratio * nadv
    s2 = torch.clamp(ratio, 1 - hp.eps_clip, 1 + hp.eps_clip) * nadv
    if hp.dual_clip:  # mappo.py:191-193
        s3 = -torch.sign(nadv) * hp.c_clip * nadv
        p_elem = -torch.max(torch.min(s1, s2), s3)
    else:
        p_elem = -torch.min(s1, s2)
    policy_loss = (p_elem * mask).sum() / msum  # mappo.py:197
    entropy_loss = -(entropy * mask).sum() / msum  # mappo.py:199
    loss = policy_loss + hp.value_loss_weight * value_loss + hp.entropy_bonus_weight * entropy_loss

    out = dict(loss=loss.detach(), policy_loss=policy_loss.detach(), value_loss=value_loss.detach(),
               entropy_loss=entropy_loss.detach())
    if want_grads:
        loss.backward()
        out.update(g_logp=new_logp.grad, g_value=values.grad, g_entropy=entropy.grad)
    bmask = mask.to(torch.bool)
    with torch.no_grad():  # mappo.py:205-217 then .mean().item() at :293-296
        stats = dict(
            advantage=torch.masked_select(adv, bmask).mean(),
            entropy=-entropy_loss.detach(),
            policy_loss=policy_loss.detach(),
            value_loss=value_loss.detach(),
            importance_weight=torch.masked_select(ratio.detach(), bmask).mean(),
            clip_ratio=torch.masked_select((s2 < s1).float(), bmask).mean(),
            value_targets=torch.masked_select(target, bmask).mean(),
        )
        if denorm_target is not None:
            stats["denorm_value"] = torch.masked_select(denorm_target, bmask).mean()
    out["stats"] = {k: float(v) for k, v in stats.items()}
    out["norm_adv"] = nadv
    return out


# --------------------------------------------------------------------------------------------
# A6: log-prob / entropy from logits (actor_critic_policy.py:303-324, :135-136)
# --------------------------------------------------------------------------------------------


def logp_entropy_from_logits_ref(logits, actions, head_sizes, available=None):
    """Sum over action heads of Categorical(logits=slice).log_prob(a_h) and .entropy().
    logits [..., sum K] (requires_grad ok), actions [..., heads] integer-valued, returns [..., 1]."""
    if available is not None:
        logits = logits.masked_fill(available == 0, -1e10)  # actor_critic_policy.py:135-136
    lps, ents, off = [], [], 0
    for h, k in enumerate(head_sizes):
        dist = torch.distributions.Categorical(logits=logits[..., off:off + k])
        lps.append(dist.log_prob(actions[..., h]))
        ents.append(dist.entropy())
        off += k
    lp = torch.stack(lps, dim=-1).sum(dim=-1, keepdim=True)
    en = torch.stack(ents, dim=-1).sum(dim=-1, keepdim=True)
    return lp, en


# --------------------------------------------------------------------------------------------
# A1: batch assembly (base/namedarray.py:588-633, base/buffer.py:118-126,
#     base/shared_memory.py:69-99)
# --------------------------------------------------------------------------------------------


def stack_leaves_ref(per_sample_leaves):
    """``recursive_aggregate(samples, lambda x: np.stack(x, axis=1))`` on a flat dict view.
    per_sample_leaves: list (one per sample) of {flat_key: ndarray[L, ...] or None}.
    None leaves are zero-filled when some other sample carries the leaf (namedarray.py:588-595);
    leaves that are None everywhere stay None (namedarray.py:630-631)."""
    keys = sorted(per_sample_leaves[0].keys())  # NamedArray iterates sorted keys (namedarray.py:282)
    out = {}
    for k in keys:
        col = [s[k] for s in per_sample_leaves]
        present = [c for c in col if c is not None]
        if not present:
            out[k] = None
            continue
        col = [np.zeros_like(present[0]) if c is None else c for c in col]
        out[k] = np.stack(col, axis=1)
    return out


def slab_get_ref(slab_leaves, slot_ids, sort=True):
    """SharedMemoryDock.get: ``buf[:, sorted(idx)]`` for every leaf (shared_memory.py:85-99)."""
    idx = np.sort(np.asarray(slot_ids)) if sort else np.asarray(slot_ids)
    return {k: (None if v is None else v[:, idx]) for k, v in slab_leaves.items()}


# --------------------------------------------------------------------------------------------
# A7: minibatch permutation (NEW capability).  Spec lives here, in numpy.
# --------------------------------------------------------------------------------------------

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3";
    same generator as cuRAND's XORWOW-free Philox4_32_10).  counter: uint32[...,4], key: uint32[...,2].
    Known-answer vectors (Random123 kat_vectors) are checked in tests/test_oracle.py."""
    c = [np.asarray(counter[..., i], dtype=np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint32).copy()
    k1 = np.asarray(key[..., 1], dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[0].astype(np.uint64)
            p1 = _M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK32).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + _W0).astype(np.uint32)
            k1 = (k1 + _W1).astype(np.uint32)
    return np.stack(c, axis=-1)


PERM_ROUNDS = 8


def _perm_round_keys(seed: int, epoch: int):
    """8 round keys = two Philox blocks keyed by the 64-bit seed, counter (block, epoch, 'SRLP', 0)."""
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    ctr = np.array([[0, epoch & 0xFFFFFFFF, 0x53524C50, 0], [1, epoch & 0xFFFFFFFF, 0x53524C50, 0]],
                   dtype=np.uint32)
    return philox4x32_10(ctr, key[None, :]).reshape(-1)  # uint32[8]


def philox_perm_ref(seed: int, epoch: int, n: int):
    """Philox-keyed bijection of [0, n): an 8-round Feistel network on ceil(log2 n) bits
    (round function = high word of a Philox-style 32x32 multiply, keyed per round), with
    cycle-walking to stay inside [0, n).  Stateless: perm[i] depends only on (seed, epoch, n, i),
    which is what lets the CUDA side evaluate it inside the gather with no index traffic.
    Returns int32[n] with perm a permutation of arange(n)."""
    if n <= 0:
        return np.zeros((0,), dtype=np.int32)
    bits = max(2, int(n - 1).bit_length())
    lb = bits // 2  # low half width
    hb = bits - lb  # high half width (>= lb)
    rk = _perm_round_keys(seed, epoch)
    lmask, hmask = np.uint32((1 << lb) - 1), np.uint32((1 << hb) - 1)

    def bij(x):
        lo = x & lmask
        hi = (x >> np.uint32(lb)) & hmask
        with np.errstate(over="ignore"):
            for r in range(PERM_ROUNDS):
                if r % 2 == 0:  # hi ^= F(lo)
                    f = ((_M0 * ((lo ^ rk[r]).astype(np.uint64))) >> np.uint64(32)).astype(np.uint32)
                    f = f ^ ((lo * np.uint32(0x9E3779B9)) >> np.uint32(16))
                    hi = (hi ^ f) & hmask
                else:  # lo ^= F(hi)
                    f = ((_M1 * ((hi ^ rk[r]).astype(np.uint64))) >> np.uint64(32)).astype(np.uint32)
                    f = f ^ ((hi * np.uint32(0xBB67AE85)) >> np.uint32(16))
                    lo = (lo ^ f) & lmask
        return (hi << np.uint32(lb)) | lo

    x = bij(np.arange(n, dtype=np.uint32))
    while True:  # cycle-walk: 2^bits < 2n so each pass fixes >= half of the stragglers on average
        bad = x >= np.uint32(n)
        if not bad.any():
            break
        x[bad] = bij(x[bad])
    return x.astype(np.int32)


def gather_lanes_ref(x, idx):
    """numpy fancy indexing on the batch axis: x[:, idx] (buffer.py:169-172 / shared_memory.py:85-99)."""
    return x[:, np.asarray(idx)]


# --------------------------------------------------------------------------------------------
# Whole hot path on CPU: the thing bench.py times as the reference arm
# --------------------------------------------------------------------------------------------


def hot_path_ref(batch: dict, hp: LossHyper, gamma: float, lmbda: float, epochs: int, minibatches: int,
                 seed: int = 0, popart: Optional[RunningMeanStdRef] = None, bootstrap_steps: int = 1,
                 burn_in_steps: int = 0, want_grads: bool = True, lanes_per_env: int = 1):
    """GAE once (mappo.py:252-257) then epochs x minibatches of loss + backward
    (mappo.py:259-274) on lane subsets chosen by ``philox_perm_ref``.

    batch: float32 CPU tensors -- reward, value, done, truncated, on_reset, old_logp: [L, N, 1];
    new_logp, v_pred, entropy: [epochs, T, N, 1] (one policy evaluation per epoch).
    Returns dict(adv, ret, per_minibatch=[...])."""
    L = batch["on_reset"].shape[0]
    adv, ret = adv_and_value_target_ref(batch["reward"], batch["value"], batch["truncated"], batch["done"],
                                        batch["on_reset"], gamma, lmbda, popart=popart)
    adv, ret = pad_last_row(adv), pad_last_row(ret)
    lo, hi = burn_in_steps, L - bootstrap_steps
    mask = 1 - batch["on_reset"][lo + 1:hi + 1]  # mappo.py:260-261
    N = mask.shape[1]
    A = lanes_per_env  # whole environments are shuffled; an env's agents stay together
    results = []
    for e in range(epochs):
        if popart is not None:
            popart.update(ret[lo:hi], mask=mask)  # mappo.py:263-264
        if minibatches > 1:
            env = philox_perm_ref(seed, e, N // A).astype(np.int64)
            perm = (env[:, None] * A + np.arange(A)[None, :]).reshape(-1)
        else:
            perm = np.arange(N)
        per = N // minibatches
        for j in range(minibatches):
            idx = torch.from_numpy(perm[j * per:(j + 1) * per].astype(np.int64))
            take = (lambda x: x) if minibatches == 1 else (lambda x: x.index_select(1, idx))
            results.append(
                ppo_loss_ref(take(batch["new_logp"][e]), take(batch["old_logp"][lo:hi]),
                             take(batch["v_pred"][e]), take(batch["value"][lo:hi]), take(ret[lo:hi]),
                             take(adv[lo:hi]), take(batch["entropy"][e]), take(mask), hp, popart=popart,
                             want_grads=want_grads))
    return dict(adv=adv, ret=ret, per_minibatch=results)
