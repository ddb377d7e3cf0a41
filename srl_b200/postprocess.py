"""Trajectory post-processor on the device: drop-in for TrajGAE (legacy/algorithm/modules/gae.py:100-139, registered as
'gae' at gae.py:142; interface api/trainer.py:84-99).

The reference walks ONE episode backwards in a python loop of small numpy operations per step.  Here the episodes a
worker has finished are laid out one after another (`[total_steps, W]` + an offset table), go to the GPU in one copy and
are all scanned by one `srl_traj_gae` launch (one thread per episode element, the arrays' own dtype, numpy's rounding
sequence -> bit-identical results).  `process(memory)` keeps the reference's one-episode signature; `process_many` is
the batched form an actor worker with many finished episodes should call.

There is no CPU path: without libsrl_b200.so / a GPU every call raises.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from srl_b200 import api, ops


def _compute_dtype(*arrays) -> np.dtype:
    """numpy's result dtype of `reward + gamma * value - value` with python-float gamma: the arrays' common dtype,
    float64 for integer / bool arrays (python float x int64 array -> float64), float16 is not supported."""
    dt = np.result_type(*[np.asarray(a).dtype for a in arrays])
    if dt.kind != "f":
        return np.dtype(np.float64)
    if dt.itemsize < 4:
        raise TypeError("TrajGAEB200: float16 rewards / values are not supported")
    return dt


class TrajGAEB200(api.TrajPostprocessor):
    """Same constructor and `process` contract as TrajGAE (gae.py:106-139)."""

    def __init__(self, gamma, lmbda, device="cuda"):
        self.gamma = gamma
        self.lmbda = lmbda
        self.device = torch.device(device)

    def process(self, memory: List):
        return self.process_many([memory])[0]

    def process_many(self, memories: Sequence[List]) -> List[List]:
        if self.device.type != "cuda":
            raise ValueError("TrajGAEB200 runs on a CUDA device only (srl_b200 has no CPU path)")
        if not memories:
            return []
        for memory in memories:
            assert np.logical_or(memory[-1].done, memory[-1].truncated).all()  # gae.py:110
        shape0 = np.asarray(memories[0][0].reward).shape
        W = int(np.prod(shape0)) if shape0 else 1
        firsts = [m[0] for m in memories]
        dt = _compute_dtype(*[f.reward for f in firsts], *[f.analyzed_result.value for f in firsts])
        lens = [len(m) for m in memories]
        offsets = np.zeros(len(memories) + 1, dtype=np.int64)
        np.cumsum(lens, out=offsets[1:])
        total = int(offsets[-1])
        reward = np.zeros((total, W), dtype=dt)
        value = np.zeros((total, W), dtype=dt)
        final_trunc = np.zeros((len(memories), W), dtype=np.uint8)
        final_has = np.zeros(len(memories), dtype=np.uint8)
        for k, memory in enumerate(memories):
            lo = int(offsets[k])
            for i, step in enumerate(memory):
                r = np.asarray(step.reward)
                if r.size != W:
                    raise ValueError(f"episode {k} step {i}: reward has {r.size} elements, the first step had {W}")
                reward[lo + i] = r.reshape(-1)
                ar = step.analyzed_result
                if ar is not None and ar.value is not None:
                    value[lo + i] = np.asarray(ar.value).reshape(-1)
                elif i != len(memory) - 1:
                    raise ValueError(f"episode {k} step {i}: only the final step may lack analyzed_result.value")
            last = memory[-1]
            if last.analyzed_result is not None and len(memory) >= 2:  # gae.py:117-123
                final_has[k] = 1
                final_trunc[k] = (np.broadcast_to(np.asarray(last.truncated).reshape(-1), (W,)) != 0)
        dev = self.device
        to_dev = lambda a: torch.from_numpy(a).to(dev, non_blocking=False)
        adv, ret = ops.traj_gae(to_dev(reward), to_dev(value), to_dev(offsets), to_dev(final_trunc), to_dev(final_has),
                                float(self.gamma), float(self.lmbda))
        adv, ret = adv.cpu().numpy(), ret.cpu().numpy()
        for k, memory in enumerate(memories):
            lo = int(offsets[k])
            for i in range(len(memory) - 1):  # the final step keeps whatever it had (gae.py:113)
                shape = np.asarray(memory[i].reward).shape
                memory[i].analyzed_result.adv = adv[lo + i].reshape(shape)
                memory[i].analyzed_result.ret = ret[lo + i].reshape(shape)
        return list(memories)


api.register_traj_postprocessor("gae_b200", TrajGAEB200)
