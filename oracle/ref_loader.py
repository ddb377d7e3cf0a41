"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* SRL reference hot-path files by path.

This module exists so that the oracle restatement in ``oracle/ref_math.py`` can be pinned
against the real reference, so that ``oracle/make_golden.py`` can generate the fixtures in
``tests/golden/``, and so that ``bench.py --impl reference`` can time the reference's own functions.
It reads ``/root/reference`` where that exists (the build container) and otherwise the byte-for-byte
copies ``oracle/stage_ref.py`` staged under ``oracle/_ref/`` (git-ignored; they travel to the GPU box
like a compiled oracle would).  Nothing under ``srl_b200/`` may import this file.

Why a loader: the reference does not import on Python 3.12 / NumPy 2:
``api/config.py:102`` (dataclass mutable defaults), ``api/policy.py:67`` (``np.bool8``),
``api/env_utils.py:2`` (``gym``).  None of those touch the arithmetic, so we pre-seed
``sys.modules`` with three tiny stand-ins and ``exec`` the real files under their real names:

    base/namedarray.py, base/gpu_utils.py, api/policy.py, api/trainer.py,
    legacy/algorithm/modules/{utils,gae,popart}.py, legacy/algorithm/ppo/mappo.py
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("SRL_REFERENCE_ROOT") or (
    "/root/reference" if os.path.isfile("/root/reference/legacy/algorithm/ppo/mappo.py") else _STAGED)


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "legacy/algorithm/ppo/mappo.py"))


def _exec_file(dotted: str, rel_path: str):
    spec = importlib.util.spec_from_file_location(dotted, os.path.join(REFERENCE_ROOT, rel_path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


def _pkg(dotted: str):
    mod = types.ModuleType(dotted)
    mod.__path__ = []  # mark as package
    sys.modules[dotted] = mod
    return mod


_LOADED = None
_USE_CUDA_PREFETCH = [False]  # set_cuda_prefetch(True): trainers built afterwards use the real PyTorchGPUPrefetcher


def set_cuda_prefetch(on: bool) -> None:
    _USE_CUDA_PREFETCH[0] = bool(on)


def load():
    """Returns a namespace with the reference's hot-path modules (cached)."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"SRL reference not found under {REFERENCE_ROOT}")

    if not hasattr(np, "bool8"):  # api/policy.py:67
        np.bool8 = np.bool_
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)  # `base.namedarray` imports normally

    # --- stand-ins for the three modules that do not import here --------------------------
    api = _pkg("api")
    cfg = types.ModuleType("api.config")

    class _Spec:
        def __init__(self, type_=None, args=None):
            self.type_ = type_
            self.args = args if args is not None else {}

    cfg.Trainer = type("Trainer", (_Spec,), {})
    cfg.Policy = type("Policy", (_Spec,), {})
    cfg.TrajPostprocessor = type("TrajPostprocessor", (_Spec,), {})
    sys.modules["api.config"] = cfg
    api.config = cfg

    env = types.ModuleType("api.environment")
    env.Action = type("Action", (), {})
    sys.modules["api.environment"] = env
    api.environment = env

    # --- the real files ---------------------------------------------------------------------
    namedarray = importlib.import_module("base.namedarray")
    _exec_file("base.gpu_utils", "base/gpu_utils.py")
    api.policy = _exec_file("api.policy", "api/policy.py")
    api.trainer = _exec_file("api.trainer", "api/trainer.py")

    legacy = _pkg("legacy")
    algo = _pkg("legacy.algorithm")
    modules = _pkg("legacy.algorithm.modules")
    legacy.algorithm = algo
    algo.modules = modules
    utils = _exec_file("legacy.algorithm.modules.utils", "legacy/algorithm/modules/utils.py")
    gae = _exec_file("legacy.algorithm.modules.gae", "legacy/algorithm/modules/gae.py")
    popart = _exec_file("legacy.algorithm.modules.popart", "legacy/algorithm/modules/popart.py")
    nstep = _exec_file("legacy.algorithm.modules.n_step_return", "legacy/algorithm/modules/n_step_return.py")
    for m in (utils, gae, popart, nstep):
        for k, v in vars(m).items():
            if not k.startswith("_"):
                setattr(modules, k, v)
    _pkg("legacy.algorithm.ppo")
    mappo = _exec_file("legacy.algorithm.ppo.mappo", "legacy/algorithm/ppo/mappo.py")

    real_prefetcher = mappo.PyTorchGPUPrefetcher

    class _HostPrefetch:
        """CPU stand-in for PyTorchGPUPrefetcher (api/trainer.py:199-228), whose constructor needs CUDA (:206): the
        same contract -- float32 copies of every leaf (:215-217), and a one-call delay: the first push() returns None,
        every later one returns the PREVIOUS sample (:219-228).  With it the unmodified MultiAgentPPO.step runs on CPU."""

        def __init__(self):
            self._next = None

        def push(self, sample):
            import torch
            nxt = (sample, namedarray.recursive_apply(sample, lambda x: torch.from_numpy(x).float()))
            cur, self._next = self._next, nxt
            return cur

    def _prefetcher():
        import torch
        return real_prefetcher() if torch.cuda.is_available() and _USE_CUDA_PREFETCH[0] else _HostPrefetch()

    mappo.PyTorchGPUPrefetcher = _prefetcher

    _LOADED = types.SimpleNamespace(namedarray=namedarray, policy=api.policy, trainer=api.trainer,
                                    utils=utils, gae=gae, popart=popart, nstep=nstep, mappo=mappo,
                                    modules=modules)
    return _LOADED


class FakePolicy:
    """Smallest object MultiAgentPPO.__init__ accepts on CPU (mappo.py:68-116)."""

    def __init__(self, popart_head=None, denormalize_value_during_rollout=False):
        import torch
        self.device = "cpu"
        self.version = 0
        self._p = torch.nn.Parameter(torch.zeros(1))
        self.popart_head = popart_head
        self.denormalize_value_during_rollout = denormalize_value_during_rollout

    def parameters(self):
        return [self._p]

    def inc_version(self):
        self.version += 1

    # popart hooks (api/policy.py; used at mappo.py:121,151,176,264)
    def normalize_value(self, x):
        return self.popart_head.normalize(x)

    def denormalize_value(self, x):
        return self.popart_head.denormalize(x)

    def update_popart(self, x, mask):
        return self.popart_head.update(x, mask=mask)
