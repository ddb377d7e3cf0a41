"""TEST INFRASTRUCTURE ONLY -- CPU restatement of `MultiAgentPPO.step` (legacy/algorithm/ppo/mappo.py:219-328)
built from oracle/ref_math.py, for end-to-end parity tests of `srl_b200.trainer.MultiAgentPPOB200`:
same policy class on CPU, same optimizer, same sample -> parameters, stats and host write-backs must agree.

`num_minibatches > 1` (absent from the reference) is defined here as: per epoch, split the
environments with `philox_perm_ref(seed, epoch, B)` and run the reference's loss on `sample[:, env_idx]`,
one optimizer step per minibatch; PopArt is updated once per epoch on the whole batch."""
from __future__ import annotations

from collections import defaultdict

import numpy as np
import torch

from oracle import ref_math as M
from srl_b200.namedarray import flatten, recursive_apply


class RefPPOTrainer:

    def __init__(self, policy, **kw):
        self.policy = policy
        self.gamma, self.lmbda = kw.get("discount_rate", 0.99), kw.get("gae_lambda", 0.97)
        self.hp = M.LossHyper(eps_clip=kw.get("eps_clip", 0.2), clip_value=kw.get("clip_value", False),
                              value_eps_clip=kw.get("value_eps_clip"), dual_clip=kw.get("dual_clip", True),
                              c_clip=kw.get("c_clip", 3), value_loss=kw.get("value_loss", "mse"),
                              value_loss_config=kw.get("value_loss_config", {}),
                              value_loss_weight=kw.get("value_loss_weight", 0.5),
                              entropy_bonus_weight=kw.get("entropy_bonus_weight", 0.01),
                              normalize_old_value=kw.get("normalize_old_value", False))
        self.burn_in, self.bootstrap = kw.get("burn_in_steps", 0), kw.get("bootstrap_steps", 1)
        self.epochs, self.minibatches = kw.get("ppo_epochs", 1), kw.get("num_minibatches", 1)
        self.popart, self.vtrace = kw.get("popart", False), kw.get("vtrace", False)
        self.max_grad_norm = kw.get("max_grad_norm")
        self.seed = kw.get("shuffle_seed", 0)
        self.recompute_adv_on_reuse = kw.get("recompute_adv_on_reuse", True)
        self.recompute_adv_among_epochs = kw.get("recompute_adv_among_epochs", False)
        self.entropy_decay_per_steps = kw.get("entropy_decay_per_steps", None)
        self.entropy_bonus_decay = kw.get("entropy_bonus_decay", 0.99)
        opt = dict(adam=torch.optim.Adam, sgd=torch.optim.SGD, rmsprop=torch.optim.RMSprop, adamw=torch.optim.AdamW)
        self.optimizer = opt[kw.get("optimizer", "adam")](policy.parameters(), **kw.get("optimizer_config", {}))
        self.frames, self.steps_done = 0, 0

    def _popart_ref(self):
        """An oracle RunningMeanStd view of the policy's head statistics (read / write through)."""
        head = self.policy.popart_head
        rms = head._PopArtValueHead__rms
        r = M.RunningMeanStdRef((1,), beta=rms._RunningMeanStd__beta, epsilon=rms._RunningMeanStd__eps)
        r.mean = rms._RunningMeanStd__mean.data.clone()
        r.mean_sq = rms._RunningMeanStd__mean_sq.data.clone()
        r.debias = rms._RunningMeanStd__debiasing_term.data.clone()
        return r

    def step(self, sample):
        if sample.truncated is None:
            sample.truncated = np.zeros_like(sample.done)
        if self.recompute_adv_on_reuse:
            sample.analyzed_result.adv = sample.analyzed_result.ret = None
        ts = recursive_apply(sample, lambda x: torch.from_numpy(x).float())  # api/trainer.py:215-217
        L = ts.on_reset.shape[0]
        lo, hi = self.burn_in, L - self.bootstrap
        B = ts.on_reset.shape[1]
        stats = defaultdict(lambda: 0.0)
        n_loss = 0
        for e in range(self.epochs):
            perm = M.philox_perm_ref(self.seed + self.steps_done, e, B) if self.minibatches > 1 else np.arange(B)
            per = B // self.minibatches
            for j in range(self.minibatches):
                env = torch.from_numpy(perm[j * per:(j + 1) * per].astype(np.int64))
                take = (lambda x: x) if self.minibatches == 1 else (lambda x: x.index_select(1, env))
                mb = recursive_apply(ts, take)
                tail = 1 if (self.vtrace and ts.analyzed_result.adv is None) else self.bootstrap
                res = self.policy.analyze(mb[:L - tail], target="ppo", burn_in_steps=self.burn_in)
                if ts.analyzed_result.adv is None:  # mappo.py:249-257 (always on the whole batch)
                    pa = self._popart_ref() if self.popart else None
                    kw = {}
                    if self.vtrace:
                        kw = dict(vtrace=True, new_logp=res.new_action_log_probs.detach(),
                                  old_logp=ts.analyzed_result.log_probs[:-1])
                    adv, ret = M.adv_and_value_target_ref(ts.reward, ts.analyzed_result.value, ts.truncated, ts.done,
                                                          ts.on_reset, self.gamma, self.lmbda, popart=pa, **kw)
                    ts.analyzed_result.adv, ts.analyzed_result.ret = M.pad_last_row(adv), M.pad_last_row(ret)
                    sample.analyzed_result.adv = ts.analyzed_result.adv.numpy()
                    sample.analyzed_result.ret = ts.analyzed_result.ret.numpy()
                    mb = recursive_apply(ts, take)
                full_mask = 1 - ts.on_reset[lo + 1:hi + 1]
                if self.popart and j == 0:  # mappo.py:263-264
                    self.policy.update_popart(ts.analyzed_result.ret[lo:hi], mask=full_mask)
                pa = self._popart_ref() if self.popart else None
                mask = take(full_mask)
                T = hi - lo
                ar = mb.analyzed_result
                nl, vp, en = res.new_action_log_probs[:T], res.state_values[:T], res.entropy[:T]
                out = M.ppo_loss_ref(nl, ar.log_probs[lo:hi], vp, ar.value[lo:hi], ar.ret[lo:hi], ar.adv[lo:hi], en, mask,
                                     self.hp, popart=pa)
                self.optimizer.zero_grad(set_to_none=True)
                torch.autograd.backward([nl, vp, en], [out["g_logp"], out["g_value"], out["g_entropy"]])
                if self.max_grad_norm is not None:
                    gn = torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                else:
                    gn = torch.sqrt(sum(p.grad.norm()**2 for p in self.policy.parameters() if p.grad is not None))
                self.optimizer.step()
                for k, v in out["stats"].items():
                    stats[k] += v
                stats["grad_norm"] += float(gn)
                n_loss += 1
            if self.recompute_adv_among_epochs:  # mappo.py:287-289
                ts.analyzed_result.adv = ts.analyzed_result.ret = None
                sample.analyzed_result.adv = sample.analyzed_result.ret = None
        for k in stats:
            stats[k] /= n_loss
        valid = slice(lo, hi)
        stats["done"] = float(ts.done[valid].mean())
        stats["truncated"] = float(ts.truncated[valid].mean())
        self.policy.inc_version()  # mappo.py:305-307
        if self.entropy_decay_per_steps and self.policy.version % self.entropy_decay_per_steps == 0:  # mappo.py:310-311
            self.hp.entropy_bonus_weight *= self.entropy_bonus_decay
        self.steps_done += 1
        self.frames += int(np.prod(sample.on_reset[valid].shape))
        info = {}
        if getattr(sample, "info_mask", None) is not None and sample.info is not None:  # mappo.py:317-324
            elapsed = sample.info_mask[valid].sum()
            if elapsed != 0:
                info = {k: float((v[valid] * sample.info_mask[valid]).sum() / elapsed) for k, v in flatten(sample.info)
                        if v is not None}
        return dict(frames=self.frames, **stats, **info), self.policy.version
