"""A10 pinned to the reference: `tests/golden/trainer_*.npz` hold what the UNMODIFIED `MultiAgentPPO.step`
(legacy/algorithm/ppo/mappo.py:219-328) returned over four calls on CPU (oracle/make_golden.py::gen_trainer: a tiny policy,
SGD, the reference's own PopArt head, the host stand-in of the prefetcher) -- stats, version, the adv / ret written back
into the host sample, the entropy coefficient and the final network parameters, for {plain, PopArt, V-trace,
bootstrap + burn-in + recompute_adv_among_epochs}.

  CPU : the oracle restatement `RefPPOTrainer` (oracle/ref_trainer.py) must reproduce them -> the restatement is pinned.
  GPU : `MultiAgentPPOB200.step` (prefetch on, so the one-call delay is part of the contract) must reproduce them.
"""
import json

import numpy as np
import pytest
import torch

from srl_b200 import api
from srl_b200.namedarray import NamedArray
from tests.doubles import TinyActorCriticPolicy
from tests.util import assert_close_ref, load_golden

CASES = ["plain", "popart", "vtrace", "boot_burn"]
OBS_DIM, NUM_ACTIONS = 6, 5


def _sample(fx, it):
    g = lambda k: fx[f"call{it}/in/{k}"].copy()
    return api.SampleBatch(obs=NamedArray(vec=g("obs_vec")), on_reset=g("on_reset"), done=g("done"), truncated=g("truncated"),
                           action=NamedArray(x=g("action_x")), reward=g("reward"),
                           info=NamedArray(episode_return=g("info_episode_return")), info_mask=g("info_mask"),
                           analyzed_result=api.AnalyzedResult(value=g("value"), log_probs=g("old_logp")))


def _policy(fx, kw, device):
    pol = TinyActorCriticPolicy(OBS_DIM, NUM_ACTIONS, device=device, popart=kw.get("popart", False), seed=3)
    sd = pol.net.state_dict()
    init = {k[len("init/"):]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("init/")}
    assert set(init) == set(sd), (sorted(init), sorted(sd))  # the double has the reference head's parameter names
    pol.net.load_state_dict(init)
    return pol


def _check_call(fx, it, stats, step, sample_prev, tol):
    assert step == int(fx[f"call{it}/step"])
    assert sorted(stats) == list(fx[f"call{it}/stat_keys"]), (sorted(stats), list(fx[f"call{it}/stat_keys"]))
    for k in stats:
        assert_close_ref(stats[k], fx[f"call{it}/stat/{k}"], tol=tol, what=f"call {it} stat {k}")
    ar = sample_prev.analyzed_result
    if bool(fx[f"call{it}/adv_is_none"]):
        assert ar.adv is None and ar.ret is None
    else:
        assert_close_ref(ar.adv, fx[f"call{it}/adv"], what=f"call {it} adv write-back")
        assert_close_ref(ar.ret, fx[f"call{it}/ret"], what=f"call {it} ret write-back")


def _check_final(fx, pol, tol):
    for k, v in pol.net.state_dict().items():
        assert_close_ref(v.cpu(), fx[f"final/{k}"], tol=tol, what=f"parameter {k} after the last call")


@pytest.mark.parametrize("name", CASES)
def test_restatement_reproduces_reference_step(name):
    """oracle/ref_trainer.py against the unmodified MultiAgentPPO.step.  The restatement has no prefetcher: call `it` of the
    reference trained on sample `it - 1`."""
    from oracle.ref_trainer import RefPPOTrainer
    fx = load_golden(f"trainer_{name}.npz")
    kw = json.loads(str(fx["kwargs_json"]))
    pol = _policy(fx, kw, "cpu")
    ref = RefPPOTrainer(pol, **kw)
    assert {k for k in fx if k.startswith("call0/stat/")} == set() and int(fx["call0/step"]) == 0  # priming call
    for it in range(1, int(fx["n_calls"])):
        s = _sample(fx, it - 1)
        stats, version = ref.step(s)
        _check_call(fx, it, stats, version, s, tol=2e-6)
        assert_close_ref(ref.hp.entropy_bonus_weight, fx[f"call{it}/entropy_bonus_weight"], tol=1e-12, what="entropy coefficient")
    _check_final(fx, pol, tol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_b200_trainer_reproduces_reference_step(name):
    """The drop-in trainer on the GPU, prefetch delay included, against the same reference-generated vectors."""
    from srl_b200.trainer import MultiAgentPPOB200
    fx = load_golden(f"trainer_{name}.npz")
    kw = json.loads(str(fx["kwargs_json"]))
    pol = _policy(fx, kw, "cuda:0")
    tr = MultiAgentPPOB200(pol, **kw)  # prefetch=True is the default, as in the reference
    samples = []
    for it in range(int(fx["n_calls"])):
        samples.append(_sample(fx, it))
        res = tr.step(samples[-1])
        if it == 0:
            assert res.stats == {} and res.step == 0  # api/trainer.py:220-223
            continue
        _check_call(fx, it, res.stats, res.step, samples[it - 1], tol=2e-5)
        assert_close_ref(tr.entropy_bonus_weight, fx[f"call{it}/entropy_bonus_weight"], tol=1e-12, what="entropy coefficient")
    _check_final(fx, pol, tol=2e-5)
