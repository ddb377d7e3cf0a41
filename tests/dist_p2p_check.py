"""Multi-GPU check of the NVLink peer-memory statistics exchange (run under torchrun on >= 2 GPUs of one node;
not collected by pytest -- the CPU suite covers the gloo path, tests/test_distributed_cpu.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_p2p_check.py

Every rank: random float64 tables; the p2p result must equal the rank-ordered sum bit for bit on every rank (and NCCL's
all-reduce within rounding), eagerly and replayed from a CUDA graph."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from srl_b200.xchg import PeerExchange  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 33 * 8
    px = PeerExchange(dist.group.WORLD, n, dev)
    g = torch.Generator(device="cpu").manual_seed(1234)
    for it in range(50):
        tables = torch.randn(world, n, dtype=torch.float64, generator=g)  # same on every rank
        mine = tables[rank].to(dev)
        out = torch.empty(n, dtype=torch.float64, device=dev)
        px.allreduce_sum(mine, out)
        expect = torch.zeros(n, dtype=torch.float64)
        for q in range(world):
            expect += tables[q]
        assert torch.equal(out.cpu(), expect), f"rank {rank} iteration {it}: p2p sum differs from the rank-ordered sum"
        ref = mine.clone()
        dist.all_reduce(ref)
        assert torch.allclose(ref.cpu(), expect, rtol=1e-14, atol=1e-14)
    # graph replay
    mine = torch.randn(n, dtype=torch.float64, device=dev)
    out = torch.empty_like(mine)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        px.allreduce_sum(mine, out)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        px.allreduce_sum(mine, out)
    for it in range(20):
        mine.copy_(torch.full((n,), float(rank + it), dtype=torch.float64))
        graph.replay()
        torch.cuda.synchronize()
        want = float(sum(q + it for q in range(world)))
        assert torch.equal(out.cpu(), torch.full((n,), want, dtype=torch.float64)), f"rank {rank} replay {it}"
    # fused: srl_group_stats_xchg sends the table from the kernel that produces it
    from srl_b200 import ops
    N, G, per = 4096, 32, 512
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    part = torch.rand(8, N, dtype=torch.float64, generator=gen).to(dev)
    idx = torch.stack([torch.randperm(N, generator=gen) for _ in range(4)]).to(torch.int32).reshape(-1).to(dev)
    ws = ops.group_stats_workspace(dev, G, True)
    local = torch.zeros((G + 1, 8), dtype=torch.float64, device=dev)
    fused = torch.zeros_like(local)
    for it in range(10):
        ops.group_stats(part, idx=idx, groups=G, per=per, out=local, whole_first=True, workspace=ws, exchange=px,
                        global_out=fused)
        ref = local.clone()
        torch.cuda.synchronize()
        dist.all_reduce(ref)
        assert torch.allclose(fused, ref, rtol=1e-13, atol=1e-13), f"rank {rank}: fused exchange differs from all_reduce"
        everyone = [torch.empty_like(fused) for _ in range(world)]
        dist.all_gather(everyone, fused)
        assert all(torch.equal(everyone[0], e) for e in everyone), "fused tables are not bit-identical across ranks"
        part.mul_(1.01)
    # latency: p2p kernel vs NCCL all-reduce, CUDA events over 200 back-to-back calls
    def timed(fn, reps=200):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3
    t_p2p = timed(lambda: px.allreduce_sum(mine, out))
    t_graph = timed(graph.replay)
    buf = mine.clone()
    t_nccl = timed(lambda: dist.all_reduce(buf))
    px.check()
    if rank == 0:
        print(f"p2p exchange ok on {world} GPUs: {t_p2p:.1f} us per call (stream launch), {t_graph:.1f} us per graph replay; "
              f"NCCL all_reduce of the same {n} doubles: {t_nccl:.1f} us", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
