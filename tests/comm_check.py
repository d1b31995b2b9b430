"""Peer-memory communicator on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 -m tests.comm_check

jck_comm_allreduce_small / jck_bn_finalize_sync / jck_bn_bwd_sums_sync against NCCL all-reduce + the unfused
kernels: hundreds of back-to-back calls of varying length (slot reuse), eager and replayed from a CUDA graph,
identical bits on every rank; then a latency comparison with a host-launched NCCL all-reduce."""
import sys

import torch
import torch.distributed as dist

from jck_generation_b200 import ops, parallel


def main():
    comm = parallel.init_from_env()
    assert comm.world_size > 1 and comm.peer is not None, "needs torchrun with >= 2 GPUs (peer communicator not opened)"
    dev = torch.device("cuda", torch.cuda.current_device())
    gen = torch.Generator(device=dev).manual_seed(100 + comm.rank)
    worst = 0.0
    for it in range(300):
        n = [1, 7, 128, 1024, 2048, 3072, 333][it % 7]
        x = torch.randn(n, device=dev, generator=gen)
        want = x.clone()
        dist.all_reduce(want)
        got = x.clone()
        ops.comm_allreduce_small(comm.peer, got)
        worst = max(worst, float((got - want).abs().max() / want.abs().max().clamp_min(1e-6)))
        # identical bits on every rank
        ref = got.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, got), f"rank {comm.rank}: all-reduce result differs from rank 0's at call {it}"
    assert worst < 1e-5, worst

    # fused SyncBN forward / backward vs unfused kernels over NCCL-reduced inputs
    C, groups, count = 256, 3, 4096.0
    stats = torch.rand(groups, 2 * C, device=dev, generator=gen) * 100
    stats[:, C:] += stats[:, :C] ** 2 / count * comm.world_size        # keep the variance positive
    gamma, beta = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
    dist.broadcast(gamma, 0); dist.broadcast(beta, 0)
    outs = []
    for fused in (False, True):
        st = stats.clone()
        rm, rv, nbt = torch.zeros(C, device=dev), torch.ones(C, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
        ss, mr = torch.empty(groups, 2 * C, device=dev), torch.empty(groups, 2 * C, device=dev)
        if fused:
            ops.bn_finalize_sync(comm.peer, st, gamma, beta, rm, rv, nbt, ss, mr, C, groups, count * comm.world_size)
        else:
            dist.all_reduce(st)
            ops.bn_finalize(st, gamma, beta, rm, rv, nbt, ss, mr, C, groups, count * comm.world_size)
        outs.append((ss, mr, rm, rv, nbt.clone()))
    for a, b in zip(*outs):
        assert torch.allclose(a.float(), b.float(), rtol=1e-5, atol=1e-6), "bn_finalize_sync != all_reduce + bn_finalize"
    sums = torch.randn(groups, 2 * C, device=dev, generator=gen)
    dg0, db0 = torch.randn(C, device=dev), torch.randn(C, device=dev)
    res = []
    for fused in (False, True):
        s, dg, db = sums.clone(), dg0.clone(), db0.clone()
        if fused:
            ops.bn_bwd_sums_sync(comm.peer, s, dg, db, C, groups, True)
        else:
            ops.bn_param_grad(s, dg, db, C, groups, True)
            dist.all_reduce(s)
        res.append((s, dg, db))
    for a, b in zip(*res):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5), "bn_bwd_sums_sync != bn_param_grad + all_reduce"

    # CUDA graph: 40 exchanges per replay, replayed 20 times
    bufs = [torch.randn(1024, device=dev, generator=gen) for _ in range(40)]
    keep = [b.clone() for b in bufs]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for b in bufs:
            ops.comm_allreduce_small(comm.peer, b)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for b in bufs:
            ops.comm_allreduce_small(comm.peer, b)
    for rep in range(20):
        for b, k in zip(bufs, keep):
            b.copy_(k)
        g.replay()
    torch.cuda.synchronize()
    for b, k in zip(bufs, keep):
        want = k.clone()
        dist.all_reduce(want)
        assert torch.allclose(b, want, rtol=1e-5, atol=1e-5), "graph replay mismatch"

    # latency: 40 small exchanges back to back
    def timed(fn, n=20):
        fn(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3 / 40
    t_p2p = timed(lambda: [ops.comm_allreduce_small(comm.peer, b) for b in bufs])
    t_graph = timed(g.replay)
    t_nccl = timed(lambda: [dist.all_reduce(b) for b in bufs])
    if comm.rank == 0:
        print(f"comm_check OK world={comm.world_size}: worst rel err {worst:.2e}; per 1024-float all-reduce: "
              f"peer-memory kernel {t_p2p:.1f} us eager / {t_graph:.1f} us in a CUDA graph, NCCL {t_nccl:.1f} us eager", flush=True)
    comm.barrier()
    torch.cuda.synchronize()
    import os
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
