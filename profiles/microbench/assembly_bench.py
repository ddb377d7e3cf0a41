"""Batch assembly: DeviceSlabBuffer vs the reference's host-side np.stack.

    python profiles/microbench/assembly_bench.py [--B 256] [--L 129]

Workload: B Atari-shaped per-environment samples (obs [L, 4, 84, 84] uint8 + the scalar leaves), the shape of SURVEY.md
App. C's probe (0.93 GB at B = 256).  Prints one JSON line: the reference path = recursive_aggregate(np.stack(axis=1))
on one host thread (base/buffer.py:118-126) followed by the prefetcher's per-leaf H2D + .float() (api/trainer.py:215-217);
ours = put() x B (pack into pinned memory + one async H2D per sample) + one srl_batch_gather; the gather kernel alone is
timed with CUDA events and set against the measured HBM bandwidth (2 x bytes: read + write)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from srl_b200 import api, ops  # noqa: E402
from srl_b200.buffer import DeviceSlabBuffer  # noqa: E402
from srl_b200.namedarray import NamedArray, flatten, recursive_aggregate, size_bytes  # noqa: E402


def sample(rng, L):
    return api.SampleBatch(
        obs=NamedArray(frame=rng.integers(0, 255, (L, 4, 84, 84), dtype=np.uint8)),
        on_reset=np.zeros((L, 1), dtype=np.uint8), done=np.zeros((L, 1), dtype=np.uint8),
        truncated=np.zeros((L, 1), dtype=np.uint8), action=NamedArray(x=rng.integers(0, 18, (L, 1)).astype(np.int32)),
        reward=rng.standard_normal((L, 1)).astype(np.float32), policy_version_steps=np.zeros((L, 1), dtype=np.int64),
        analyzed_result=api.AnalyzedResult(value=rng.standard_normal((L, 1)).astype(np.float32),
                                           log_probs=-rng.random((L, 1)).astype(np.float32)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=256)
    ap.add_argument("--L", type=int, default=129)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--copy-threads", type=int, default=1, help="host threads per large-leaf copy into pinned memory")
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    samples = [sample(rng, a.L) for _ in range(a.B)]
    nbytes = sum(size_bytes(s) for s in samples)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0

    # ---- reference path on the host: np.stack, then per-leaf H2D + .float() --------------------------------------
    t_stack, t_h2d = [], []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        batch = recursive_aggregate(samples, lambda xs: np.stack(xs, axis=1))
        t1 = time.perf_counter()
        dev = [torch.from_numpy(v).cuda().float() for _, v in flatten(batch) if v is not None]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_stack.append(t1 - t0)
        t_h2d.append(t2 - t1)
        del dev
    # ---- ours -----------------------------------------------------------------------------------------------------
    buf = DeviceSlabBuffer(max_size=2, reuses=1, batch_size=a.B, copy_threads=a.copy_threads)
    t_put, t_total = [], []
    for _ in range(a.reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in samples:
            buf.put(s)
        t1 = time.perf_counter()
        entry = buf.get()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_put.append(t1 - t0)
        t_total.append(t2 - t0)
    got = dict(flatten(entry.sample))
    want = recursive_aggregate(samples, lambda xs: np.stack(xs, axis=1))
    for k, v in flatten(want):
        if v is not None and k != "trainer_worker_recv_timestamp":
            assert np.array_equal(got[k].cpu().numpy(), v), k
    # ---- the gather kernel alone ----------------------------------------------------------------------------------
    frame = got["obs.frame"]
    src = torch.empty((a.B,) + (a.L,) + tuple(frame.shape[2:]), dtype=torch.uint8, device="cuda")  # [slots, L, row]
    dst = torch.empty_like(frame)
    row = int(np.prod(frame.shape[2:]))
    from srl_b200._lib import LeafDesc
    desc = [LeafDesc(src.data_ptr(), dst.data_ptr(), row, a.B, row, a.L * row)]
    idx = torch.arange(a.B, dtype=torch.int32, device="cuda")
    ev = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.stack_samples(desc, idx, a.L, a.B)
        e1.record()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    k_ms = sorted(x.elapsed_time(y) for x, y in ev)[len(ev) // 2]
    k_gbs = 2 * frame.numel() / k_ms / 1e6
    best = lambda xs: min(xs[1:] if len(xs) > 1 else xs)
    line = dict(
        workload=f"{a.B} Atari-shaped samples, L={a.L}, {nbytes / 1e9:.2f} GB",
        reference=dict(np_stack_s=best(t_stack), np_stack_gbs=nbytes / best(t_stack) / 1e9, h2d_float_s=best(t_h2d),
                       total_s=best(t_stack) + best(t_h2d), threads=1,
                       what="recursive_aggregate(np.stack(axis=1)) + per-leaf pageable H2D + .float() (base/buffer.py:118-126, api/trainer.py:215-217)"),
        ours=dict(put_s=best(t_put), put_gbs=nbytes / best(t_put) / 1e9, total_s=best(t_total),
                  total_gbs=nbytes / best(t_total) / 1e9,
                  what="DeviceSlabBuffer: B x (pack into pinned + one async H2D) + one srl_batch_gather; bit-exact vs np.stack"),
        gather_kernel=dict(ms=k_ms, gbs=k_gbs, frac_of_measured_hbm=k_gbs / peak, bytes=2 * frame.numel(),
                           what="srl_batch_gather on the frame leaf alone, [slots, L, row] -> [L, B, row], CUDA events"),
        speedup_total=(best(t_stack) + best(t_h2d)) / best(t_total))
    print(json.dumps(line))


if __name__ == "__main__":
    main()
