#!/usr/bin/env python
"""bench.py -- GAE + PPO-loss transitions/s on the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2_atari_large] [--impl reference]

A "step" = one pass of the hot path over one synthetic batch: GAE once, then
(epochs x minibatches) fused loss forward+backward covering every transition once per epoch.
`value` = T*N transitions per step (x ranks) / device time with inputs resident in HBM; `e2e` = the same
through `HotPath.run_host` with pinned HOST buffers in and out (H2D + D2H inside the timed region).
N > 1 (torchrun): every rank owns its own slice of environments (weak scaling, the full config per GPU);
the only data-path exchange is the SUM of the float64 statistics table (NVLink peer-memory mailboxes, or NCCL).
`--impl reference` times the reference's CPU path (the oracle port: /root/reference is a Python tree that
does not exist on the GPU box) on the host cores, same config / metric / unit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from srl_b200 import synth  # noqa: E402

METRIC = "gae_ppo_loss_transitions_per_sec"
UNIT = "transitions/s"
GAE_BYTES = 19  # 11 B read + 8 B written per scanned row-lane
LOSS_BYTES = {True: 41, False: 37}  # per transition per pass, with / without value clipping


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def workload_name(cfg):
    return (f"{cfg.name}: T={cfg.T}, B={cfg.B} envs, A={cfg.A} agents, L={cfg.L} rows, {cfg.epochs} epochs x "
            f"{cfg.minibatches} minibatches, clip_value={cfg.clip_value}, dual_clip={cfg.dual_clip}, "
            f"value_loss={cfg.value_loss}, popart={cfg.popart}")


def hyper_kwargs(cfg):
    return dict(eps_clip=cfg.eps_clip, clip_value=cfg.clip_value, dual_clip=cfg.dual_clip, c_clip=cfg.c_clip,
                value_loss=cfg.value_loss, value_loss_weight=cfg.value_loss_weight,
                entropy_bonus_weight=cfg.entropy_bonus_weight,
                value_loss_config=({"delta": cfg.value_loss_delta} if cfg.value_loss == "huber" else
                                   {"beta": cfg.value_loss_delta} if cfg.value_loss == "smoothl1" else None))


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, steps, warmup, budget_s=None):
    """Times oracle.ref_math.hot_path_ref (GAE in float64 with the reference's python scan, then
    epochs x minibatches of loss + autograd backward) on all host threads, full config per step."""
    from oracle import ref_math as M  # checker / baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    s = synth.make_sample_scalars(cfg, seed=0)
    pol = synth.make_policy_outputs(cfg, s, seed=1)
    batch = {k: torch.from_numpy(v).float() for k, v in s.items()}  # the prefetcher's .float() (api/trainer.py:217)
    batch.update({k: torch.from_numpy(v) for k, v in pol.items()})
    hp = M.LossHyper(**hyper_kwargs(cfg))
    pa = M.RunningMeanStdRef((1,)) if cfg.popart else None

    def one():
        M.hot_path_ref(batch, hp, cfg.gamma, cfg.lmbda, cfg.epochs, cfg.minibatches, seed=0, popart=pa,
                       bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps)

    for _ in range(warmup):
        one()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s and len(times) >= 3:
            break
    mean = sum(times) / len(times)
    return dict(value=cfg.transitions / mean, ms_per_step=mean * 1e3, steps=len(times),
                cores=torch.get_num_threads(), best_ms=min(times) * 1e3)


def run_reference_arm(args, cfg, rank, world):
    if rank != 0:
        return  # rank 0 alone runs the CPU arm; the others exit 0 without work
    r = cpu_reference_run(cfg, args.steps, args.warmup)
    sample = (f"{r['steps']} full steps of {workload_name(cfg)} on {r['cores']} host threads "
              f"(oracle port of mappo.py:118-217 + gae.py:8-97, torch-CPU, float64 GAE)")
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps"], warmup=args.warmup,
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference", config=dict(workload=workload_name(cfg)),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=sample),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# host placement
# ----------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(device_index):
    """Runs this process on the CPUs NVML reports as local to the GPU, so that the pinned staging buffers are
    allocated on that NUMA node and the e2e copies do not cross the socket interconnect (box to box the same bench
    measured 0.72 and 1.27 ms per e2e step without it).  Returns the previous affinity (restored for the CPU arm)."""
    try:
        import pynvml
        before = os.sched_getaffinity(0)
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID((uuid if uuid.startswith("GPU-") else "GPU-" + uuid).encode())
        words = (max(before) + 64) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {i for i in before if (mask[i // 64] >> (i % 64)) & 1}
        if cpus:
            os.sched_setaffinity(0, cpus)
        return before
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.rows, self.proc = [], None
        try:
            uuid = str(torch.cuda.get_device_properties(device_index).uuid)
            ident = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        except Exception:
            ident = str(device_index)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ident, f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons, loaded = [], [], set(), []
        for ts, line in self.rows:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                c, m, util = float(p[0]), float(p[1]), float(p[3])
            except ValueError:
                continue
            inside = t0 is None or (t0 - 0.05 <= ts <= t1 + 0.05)
            if not inside:
                continue
            sm.append(c)
            mx.append(m)
            if util > 0:
                loaded.append(c)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pick = loaded or sm
        return dict(sm_mhz=statistics.median(pick) if pick else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), samples_under_load=len(loaded))


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def event_ms(pairs):
    return [a.elapsed_time(b) for a, b in pairs]


def run_ours(args, cfg, rank, world, local_rank):
    import torch.distributed as dist
    from srl_b200 import ops
    from srl_b200.hotpath import HotPath

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity_before = bind_to_gpu_numa(local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    ops._lib.load_library()  # fail loudly if the CUDA library is missing

    # ---- synthetic batch for this rank's slice of environments (weak scaling) ----------------------
    s = synth.make_sample_scalars(cfg, seed=1000 * rank)
    pol = synth.make_policy_outputs(cfg, s, seed=1000 * rank + 1)
    E, Mb, T, N = cfg.epochs, cfg.minibatches, cfg.T, cfg.N
    hp = HotPath(cfg.L, cfg.B, cfg.A, gamma=cfg.gamma, lmbda=cfg.lmbda, hyper=ops.LossHyper(**hyper_kwargs(cfg)),
                 bootstrap_steps=cfg.bootstrap_steps, burn_in_steps=cfg.burn_in_steps, epochs=E, minibatches=Mb, seed=0,
                 popart=cfg.popart, device=dev, process_group=pg, fuse_gather=not args.explicit_gather,
                 graph_branches=args.branches, shuffle_block=args.shuffle_block, use_pack=not args.no_pack,
                 batch_losses=not args.no_batch, stats_exchange=args.exchange, fuse_stats=not args.no_fuse_stats)
    n = hp.n_mb
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v.reshape(cfg.L, N))).pin_memory() for k, v in s.items()}
    hp.load_sample(pinned)
    # policy outputs as the policy would emit them: one contiguous [T, n] block per (epoch, minibatch)
    blk = args.shuffle_block
    perm = [ops.philox_perm(0, e, cfg.B // blk, cfg.A * blk).long() if Mb > 1 else None for e in range(E)]
    pol_all = torch.empty((E, Mb, 3, T, n), dtype=torch.float32, device=dev)
    for e in range(E):
        for q, k in enumerate(("new_logp", "v_pred", "entropy")):
            full = torch.from_numpy(pol[k][e].reshape(T, N)).to(dev)
            for j in range(Mb):
                pol_all[e, j, q] = full if Mb == 1 else full.index_select(1, perm[e][j * n:(j + 1) * n])
    pol_dev = [[tuple(pol_all[e, j, q] for q in range(3)) for j in range(Mb)] for e in range(E)]
    pol_host = pol_all.cpu().pin_memory()
    out_host = dict(adv=torch.empty((cfg.L, N), dtype=torch.float32).pin_memory(),
                    ret=torch.empty((cfg.L, N), dtype=torch.float32).pin_memory(),
                    out=torch.empty((E * Mb, 16), dtype=torch.float64).pin_memory())
    # L2 flush between timed iterations: write a 256 MiB buffer (> 126 MB L2), then READ a second one.  The write
    # alone leaves the L2 full of dirty lines whose write-back the timed kernels would then pay for (it showed as
    # ~20 us on a 65 us kernel); the read pass evicts them and leaves clean lines, so the step starts cold.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(64 << 20, dtype=torch.int32, device=dev)  # 256 MiB

    def flush_l2():
        flush.zero_()
        flush_rd.max()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_l2_fn = flush_l2

    def timed_steps(k, fn, flush_l2=True):
        pairs = []
        for _ in range(k):
            if flush_l2:
                flush_l2_fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            pairs.append((a, b))
        torch.cuda.synchronize()
        return event_ms(pairs)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    step = lambda: hp.run_device(pol_dev, use_graph=True)
    for _ in range(max(args.warmup, 3)):
        flush_l2()
        step()
    barrier()
    t_region0 = time.perf_counter()
    ms = timed_steps(args.steps, step)
    barrier()
    total_ms = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = cfg.transitions * world / (ms_per_step * 1e-3)

    # ---- warm-L2 variant (no flush), informational ------------------------------------------------
    ms_warm = timed_steps(min(args.steps, 200), step, flush_l2=False)

    # ---- e2e: pinned host buffers in / out through HotPath.run_host -------------------------------
    e2e_steps = max(3, min(args.steps, args.e2e_steps))
    for _ in range(3):
        hp.run_host(pinned, pol_host, out_host)
    barrier()
    wall = []
    nbytes = None
    for _ in range(e2e_steps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = hp.run_host(pinned, pol_host, out_host)
        wall.append(time.perf_counter() - t0)
    e2e_t = torch.tensor([sum(wall)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = cfg.transitions * world / (e2e_t.item() / e2e_steps)

    # ---- per-kernel durations for the roofline: CUDA events around graphs of one kernel kind ------
    peak, peak_src = peaks()
    k_reps = 20

    def graph_of(fn):
        s_ = torch.cuda.Stream()
        s_.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s_):
            fn()
        torch.cuda.current_stream().wait_stream(s_)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    pg_saved, hp.pg = hp.pg, None
    # the loss launches exactly as the step issues them (one batched launch, one per epoch with PopArt, or one per
    # minibatch with --no-batch); PopArt's tiny update kernels ride along in the per-epoch case
    loss_launches = (E if cfg.popart else -(-(E * Mb) // 32)) if hp._immediate else E * Mb
    # stand-alone there is no kernel ahead to overlap with: launched without the programmatic attribute (the loss kernel
    # then also skips the L2 requests it only issues when it can become resident under the scan)
    pdl_was = ops.set_pdl(False)
    g_loss = graph_of(lambda: hp._run_losses(pol_dev))
    lf = hp.leaf
    pack_kw = dict(old_logp=lf["old_logp"], pack=hp.pack, lane_aos=hp.lane_aos) if hp.pack is not None else {}
    g_gae = graph_of(lambda: ops.gae_scan(lf["reward"], lf["value"], lf["done"], lf["truncated"], lf["on_reset"], cfg.gamma,
                                          cfg.lmbda, row_lo=hp.row_lo, row_hi=hp.row_hi,
                                          popart_mean_std=hp.popart_mean_std(), adv=hp.adv, ret=hp.ret,
                                          lane_part=hp.lane_part, **pack_kw))
    ops.set_pdl(pdl_was)
    hp.pg = pg_saved
    loss_ms = statistics.mean(timed_steps(k_reps, g_loss.replay)) / loss_launches
    gae_ms = statistics.mean(timed_steps(k_reps, g_gae.replay))
    loss_ms_warm = statistics.mean(timed_steps(k_reps, g_loss.replay, flush_l2=False)) / loss_launches
    gae_ms_warm = statistics.mean(timed_steps(k_reps, g_gae.replay, flush_l2=False))
    loss_bytes = LOSS_BYTES[bool(cfg.clip_value)] * T * n * (E * Mb // loss_launches)
    gae_bytes = GAE_BYTES * (cfg.L - 1) * N
    kern = {
        "ppo_loss_kernel": dict(launches_per_step=loss_launches, ms_per_launch=loss_ms, bytes_per_launch=loss_bytes,
                                gbs=loss_bytes / loss_ms / 1e6, gbs_l2_warm=loss_bytes / loss_ms_warm / 1e6,
                                step_share=loss_ms * loss_launches / ms_per_step),
        "gae_scan_kernel": dict(launches_per_step=1, ms_per_launch=gae_ms, bytes_per_launch=gae_bytes,
                                gbs=gae_bytes / gae_ms / 1e6, gbs_l2_warm=gae_bytes / gae_ms_warm / 1e6,
                                step_share=gae_ms / ms_per_step),
    }
    dom = max(kern, key=lambda k_: kern[k_]["step_share"])
    traffic = None  # DRAM bytes per launch of the dominant kernel from the committed ncu capture of this workload
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get(cfg.name, {}).get(dom)
    except Exception:
        pass
    roofline = dict(bound="hbm", kernel=dom, achieved=kern[dom]["gbs"], peak=peak, unit="GB/s",
                    frac=kern[dom]["gbs"] / peak, traffic=traffic, peak_source=peak_src,
                    algorithmic_bytes_per_launch=kern[dom]["bytes_per_launch"],
                    ms_per_launch=kern[dom]["ms_per_launch"])
    step_bytes = gae_bytes + LOSS_BYTES[bool(cfg.clip_value)] * T * N * E
    t_region1 = time.perf_counter()
    clocks = sampler.stop(t_region0, t_region1) if sampler else None

    # ---- CPU baseline beside it (rank 0, N = 1 only) ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if affinity_before:
            os.sched_setaffinity(0, affinity_before)  # the CPU arm gets every host core
        r = cpu_reference_run(cfg, steps=200, warmup=2, budget_s=args.cpu_budget_s)
        cpu = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port",
                   sample=f"{r['steps']} full steps ({r['ms_per_step']:.1f} ms each) of the same workload, oracle port "
                          f"(torch-CPU, float64 GAE python scan + loss + autograd backward), {r['cores']} threads")
    if rank == 0:
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
            data="synthetic",
            config=dict(workload=workload_name(cfg), transitions_per_step_per_gpu=cfg.transitions,
                        l2="flushed between timed iterations (256 MiB device write, then 256 MiB device read, before each step; working set "
                           f"{step_bytes / 1e6:.0f} MB algorithmic)",
                        timing="CUDA events per step on the launch stream, sum over steps, max over ranks",
                        launch="one CUDA graph per step" + ("" if world == 1 else (
                            f" ({hp.exchange_kind} exchange of the float64 stats table captured inside)"
                            if hp._graph_a is None else f" split in two around the {hp.exchange_kind} exchange of the float64 stats table")),
                        loss_launch=("batched: %d launch(es) per step covering %d minibatches each" %
                                     (loss_launches, E * Mb // loss_launches)) if hp._immediate else "one launch per minibatch",
                        sample_side="K2 pack (float4 per transition)" if hp.pack is not None else "separate leaves",
                        minibatch_stats="added inside the loss kernel from K2's per-lane sums (table on a side branch)"
                        if hp.fuse_stats else "srl_group_stats table between K2 and the loss",
                        minibatch_gather=("fused into the loss loads (lane_idx)" if hp.fuse_gather else "explicit K5 gather") +
                        f", Philox permutation of {hp.shuffle_block}-environment blocks",
                        graph_branches=hp.graph_branches,
                        programmatic_dependent_launch=bool(ops.pdl_enabled()),
                        kernel_timing="kernels[*]: each kernel alone in a graph, launched without the programmatic "
                                      "attribute, CUDA events, L2 flushed"),
            clocks=clocks,
            e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=nbytes["h2d_bytes"],
                     d2h_bytes_per_step=nbytes["d2h_bytes"], ms_per_step=e2e_t.item() / e2e_steps * 1e3, steps=e2e_steps,
                     what="HotPath.run_host: pinned host sample + policy outputs -> H2D -> step -> D2H of adv, ret (the "
                          "reference's host mirror, mappo.py:254-257) and the loss/stats table; gradients stay in HBM for "
                          "the policy's backward; one CUDA graph; wall clock incl. final stream sync"),
            gpu_launches=hp.count_launches() * args.steps,
            gpu_launches_per_step=hp.count_launches(),
            roofline=roofline, kernels=kern,
            step=dict(algorithmic_bytes=step_bytes, gbs=step_bytes / ms_per_step / 1e6,
                      frac_of_peak=step_bytes / ms_per_step / 1e6 / peak,
                      ms_per_step_l2_warm=statistics.mean(ms_warm),
                      value_l2_warm=cfg.transitions * world / (statistics.mean(ms_warm) * 1e-3)),
            cpu_baseline=cpu)
        print(json.dumps(line), flush=True)
    del hp, g_loss, g_gae  # CUDA graphs and peer mailboxes go before the process group does
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2_atari_large", choices=sorted(synth.CONFIGS))
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shuffle-block", type=int, default=1,
                    help="environments per shuffled block (1 = per-environment permutation; 8 = one 32-byte sector)")
    ap.add_argument("--branches", type=int, default=16, help="parallel CUDA-graph branches for the per-minibatch launches")
    ap.add_argument("--explicit-gather", action="store_true", help="separate K5 gather launch instead of gather-on-load")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="statistics exchange across ranks: NVLink peer-memory mailboxes (p2p) or NCCL all-reduce")
    ap.add_argument("--no-fuse-stats", action="store_true",
                    help="A/B: minibatch statistics from the srl_group_stats table instead of inside the loss kernel")
    ap.add_argument("--no-pack", action="store_true", help="A/B: gather the five sample leaves instead of K2's pack")
    ap.add_argument("--no-batch", action="store_true", help="A/B: one loss launch per minibatch (parallel graph branches)")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 50:
            args.steps = 50  # a CPU step of cfg2 is ~0.2 s; keep the whole arm within minutes
        run_reference_arm(args, cfg, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, cfg, rank, world, local_rank)


if __name__ == "__main__":
    main()
