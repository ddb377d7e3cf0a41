"""Multi-GPU correctness on the driver's path: when the box shows >= 2 GPUs, spawn one rank per GPU (torchrun, NCCL) and
check (1) the NVLink peer-memory exchange against the rank-ordered float64 sum, bit for bit on every rank, eagerly, from a
CUDA graph and fused into the statistics kernel (tests/dist_p2p_check.py), and (2) the whole step at 2 ranks through
`bench.py`'s parity_check (exchanged table bit-exact, one minibatch's loss / gradients against the oracle with the GLOBAL
sums).  Reference pattern: legacy/tests/modules_test.py:273-299 (two ranks compare their shards with the whole batch).
On a one-GPU box both tests skip; the gloo twins run in tests/test_distributed_cpu.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

from tests.util import ROOT

pytestmark = pytest.mark.gpu


def _ranks():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip(f"needs >= 2 visible GPUs, this box shows {n}")
    return 2 if n < 4 else (4 if n < 8 else 8)


def _torchrun(n, script, *args, port=29541, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), script, *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_peer_exchange_bit_exact_on_every_rank():
    n = _ranks()
    r = _torchrun(n, os.path.join(ROOT, "tests", "dist_p2p_check.py"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"p2p exchange ok on {n} GPUs" in r.stdout, r.stdout[-2000:]
    print(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("config,scaling", [("cfg2_atari_large", "weak"), ("cfg4_football_11v11", "strong")])
def test_step_parity_across_ranks(config, scaling):
    n = _ranks()
    r = _torchrun(n, os.path.join(ROOT, "bench.py"), "--gpus", str(n), "--config", config, "--scaling", scaling, "--steps", "20",
                  "--warmup", "3", "--e2e-steps", "3", "--no-cpu-baseline", "--no-extras", port=29543)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    pc = line["parity_check"]
    assert line["n_gpus"] == n and pc["ranks"] == n
    assert pc["stats_table_bit_exact_on_every_rank"] is True, pc
    assert pc["ok"] and pc["loss_err"] <= 1e-5 and pc["grad_err"] <= 1e-5, pc
