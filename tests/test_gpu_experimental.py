"""GPU checks of code paths that are built but NOT yet measured or enabled by default (tuning knobs read once from the
environment, so each check runs in a subprocess).  Opt-in: SRL_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu

  SRL_LOSS_LANES=2   pack-form loss kernel with two lanes per thread (csrc/ppo_loss_pack2.cu): must equal the default
                     four-lane kernel to the last bit of every gradient and within the fp32 block-sum tolerance of the sums
                     (profiles/r1d_notes.md, candidate 1 for the next round).
"""
import os
import subprocess
import sys

import pytest

from tests.util import ROOT

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SRL_TEST_EXPERIMENTAL") != "1", reason="opt-in: SRL_TEST_EXPERIMENTAL=1")]

_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from srl_b200 import ops, synth
cfg = synth.PathConfig("x", T=24, B=2048, epochs=2, minibatches=4, p_end=0.02, clip_value=True, dual_clip=False,
                       value_loss="huber", value_loss_delta=10.0, value_loss_weight=1.0)
s = synth.make_sample_scalars(cfg, 0)
pol = synth.make_policy_outputs(cfg, s, 1)
flat = lambda x: np.ascontiguousarray(x.reshape(x.shape[0], -1))
d = {k: torch.from_numpy(flat(v)).cuda() for k, v in s.items()}
T, N, n = cfg.T, cfg.N, cfg.N // 4
pack = torch.empty((cfg.L, N, 4), dtype=torch.float32, device="cuda")
adv, ret, part = ops.gae_scan(d["reward"], d["value"], d["done"], d["truncated"], d["on_reset"], cfg.gamma, cfg.lmbda,
                              row_lo=0, row_hi=T, old_logp=d["old_logp"], pack=pack)
perm = ops.philox_perm(5, 0, N, n_epochs=2)
hp = ops.LossHyper(clip_value=True, dual_clip=False, value_loss="huber", value_loss_config=dict(delta=10.0), value_loss_weight=1.0)
ws = ops.new_loss_workspace("cuda", 8)
probs = []
for e in range(2):
    for j in range(4):
        idx = perm[e, j * n:(j + 1) * n].contiguous()
        sel = lambda k: torch.from_numpy(flat(pol[k][e])).cuda().index_select(1, idx.long()).contiguous()
        probs.append(dict(new_logp=sel("new_logp"), v_pred=sel("v_pred"), entropy=sel("entropy"), lane_idx=idx,
                          grads=tuple(torch.empty((T, n), device="cuda") for _ in range(3)), workspace=ws[len(probs)],
                          out=torch.empty(16, dtype=torch.float64, device="cuda"),
                          out_f32=torch.empty(4, device="cuda")))
ops.ppo_loss_batched(probs, None, None, None, None, None, hp, pack=pack[:T], lane_part=part)
torch.cuda.synchronize()
np.savez(sys.argv[1], **{f"g{k}_{i}": g.cpu().numpy() for k, q in enumerate(probs) for i, g in enumerate(q["grads"])},
         **{f"out{k}": q["out"].cpu().numpy() for k, q in enumerate(probs)})
""" % ROOT


def _run(tmp_path, tag, env):
    out = tmp_path / f"{tag}.npz"
    subprocess.run([sys.executable, "-c", _SCRIPT, str(out)], check=True, env=dict(os.environ, **env), timeout=300)
    import numpy as np
    with np.load(out) as z:
        return {k: z[k] for k in z.files}


def test_two_lane_loss_kernel_equals_four_lane_kernel(tmp_path):
    import numpy as np
    a = _run(tmp_path, "lanes4", {"SRL_LOSS_LANES": "4"})
    b = _run(tmp_path, "lanes2", {"SRL_LOSS_LANES": "2"})
    assert sorted(a) == sorted(b)
    for k in a:
        if k.startswith("g"):
            assert np.array_equal(a[k], b[k]), k  # per-element math is identical
        else:
            np.testing.assert_allclose(b[k], a[k], rtol=2e-6, atol=1e-9, err_msg=k)  # fp32 block sums, another tiling
