"""world_size-2 (gloo, CPU) checks of the data-parallel host logic: row sharding, flat parameter
buckets, and the SyncBN / gradient-averaging conventions the engines rely on
(jck_generation_b200/parallel.py, engine.py).  The CUDA kernels themselves are covered by the -m gpu
tests; here the per-rank arithmetic is restated in torch so the *collective plumbing* can be proven:
N ranks on B/N rows each == one process on B rows == torch autograd on the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _syncbn_conv_rank(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from jck_generation_b200 import parallel
    comm = parallel.init_from_env(backend="gloo")
    assert isinstance(comm, parallel.TorchComm) and comm.world_size == world and comm.rank == rank

    gen = torch.Generator().manual_seed(0)
    B, Cin, C, H = 8, 3, 16, 8
    x_all = torch.randn(B, Cin, 2 * H, 2 * H, generator=gen)
    w = torch.randn(C, Cin, 4, 4, generator=gen) * 0.1
    gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.1
    up_all = torch.randn(B, C, H, H, generator=gen)           # d(loss_sum)/d(act), per sample

    x = parallel.shard_rows(x_all, comm)
    # loss = mean over the GLOBAL batch (engine convention: jck_head_bwd mean_count = rows x world), so the ranks' parameter
    # gradients SUM to the reference's global-batch gradient
    up = parallel.shard_rows(up_all, comm) / (x.shape[0] * comm.world_size)
    # ---- forward, as engine.trunk_forward does it
    y = F.conv2d(x, w, stride=2, padding=1)
    stats = torch.cat([y.sum((0, 2, 3)), (y * y).sum((0, 2, 3))])
    comm.allreduce_sum_(stats)
    count = y.shape[0] * H * H * comm.world_size
    mean = stats[:C] / count
    var = stats[C:] / count - mean * mean
    rstd = (var + 1e-5).rsqrt()
    xhat = (y - mean.view(1, C, 1, 1)) * rstd.view(1, C, 1, 1)
    pre = xhat * gamma.view(1, C, 1, 1) + beta.view(1, C, 1, 1)
    # ---- backward, as engine.trunk_backward does it
    g = torch.where(pre > 0, up, 0.2 * up)
    sums = torch.cat([g.sum((0, 2, 3)), (g * xhat).sum((0, 2, 3))])
    dgamma_local, dbeta_local = sums[C:].clone(), sums[:C].clone()      # BEFORE the all-reduce
    comm.allreduce_sum_(sums)
    dy = (gamma * rstd).view(1, C, 1, 1) * (g - sums[:C].view(1, C, 1, 1) / count - xhat * sums[C:].view(1, C, 1, 1) / count)
    dw_local = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=2, padding=1)

    # ---- flat buffer (reverse parameter order) + bucketed gradient exchange, started as the gradients become final
    mod = torch.nn.Module()
    mod.w = torch.nn.Parameter(w.clone()); mod.g = torch.nn.Parameter(gamma.clone()); mod.b = torch.nn.Parameter(beta.clone())
    flat = parallel.FlatParams(mod)
    A = flat.ALIGN                                                    # every parameter starts on a 128-byte boundary
    assert flat.layout == [2, 1, 0] and flat.offsets[2][0] == 0 and flat.offsets[1][0] == A and flat.offsets[0][0] == 2 * A
    sync = parallel.GradBuckets(flat, comm, min_elems=2 * A)          # buckets: [b, g] and [w]
    assert [(lo, hi) for lo, hi, _ in sync.buckets] == [(0, 2 * A), (2 * A, flat.numel)]
    sync.begin()
    mod.g.grad.copy_(dgamma_local); mod.b.grad.copy_(dbeta_local)
    sync.ready(mod.g, mod.b)                                          # first bucket complete: its all-reduce starts here
    assert sync.issued == [True, False]
    mod.w.grad.copy_(dw_local)
    sync.ready(mod.w)
    sync.finish()
    assert sync.issued == [True, True] and not sync.handles
    mod.zero_grad()            # sets .grad = None; rebind must re-attach the bucket views
    flat.rebind()
    assert mod.w.grad.data_ptr() == flat.grad[2 * A:].data_ptr() and mod.b.grad.data_ptr() == flat.grad.data_ptr()

    # ---- single-process truth on the full batch: loss = mean over the GLOBAL batch
    wt, gt, bt = w.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yt = F.conv2d(x_all, wt, stride=2, padding=1)
    at = F.leaky_relu(F.batch_norm(yt, None, None, gt, bt, True, 0.1, 1e-5), 0.2)
    (at * up_all).sum().div(B).backward()
    per = B // world
    ok = (torch.allclose(F.leaky_relu(pre, 0.2), at[rank * per:(rank + 1) * per].detach(), atol=1e-5)
          and torch.allclose(mod.w.grad, wt.grad, atol=1e-5, rtol=1e-4)
          and torch.allclose(mod.g.grad, gt.grad, atol=1e-5, rtol=1e-4)
          and torch.allclose(mod.b.grad, bt.grad, atol=1e-5, rtol=1e-4))
    ret[rank] = bool(ok)
    comm.barrier()
    dist.destroy_process_group()


def test_syncbn_and_gradient_exchange_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_syncbn_conv_rank, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_local_comm_and_sharding():
    from jck_generation_b200 import parallel
    c = parallel.LocalComm()
    t = torch.arange(8.).view(4, 2)
    assert parallel.shard_rows(t, c) is not None and torch.equal(parallel.shard_rows(t, c), t)
    assert c.allreduce_sum_(t) is t and c.allreduce_mean_(t) is t
    os.environ.pop("WORLD_SIZE", None)
    assert isinstance(parallel.init_from_env(), parallel.LocalComm)


def test_flat_params_keep_module_semantics():
    from jck_generation_b200 import parallel
    lin = torch.nn.Linear(4, 3)
    before = {k: v.clone() for k, v in lin.state_dict().items()}
    flat = parallel.FlatParams(lin)
    assert flat.numel == 2 * flat.ALIGN            # weight (12) and bias (3), each padded to a 128-byte boundary
    for k, v in lin.state_dict().items():
        assert torch.equal(v, before[k])
    flat.flat.mul_(2)                      # the parameters are views of the bucket
    assert torch.equal(lin.weight.detach(), before["weight"] * 2)
    lin.load_state_dict(before)            # in-place copy keeps the views alive
    assert torch.equal(flat.views(flat.flat)[0], before["weight"])


def test_rank_rows_of_a_global_batch():
    """data parallel loaders: the batch size is the GLOBAL batch, rank r owns rows [r B/W, (r+1) B/W); the ranks'
    rows tile the global batch without overlap, a remainder is dropped on every rank"""
    import torch
    from jck_generation_b200 import parallel
    from jck_generation_b200.preprocess.synthetic import SyntheticLoader
    for n, world in ((512, 8), (128, 4), (37, 2), (5, 8)):
        rows = [list(range(n))[parallel.local_slice(n, r, world)] for r in range(world)]
        assert all(len(x) == n // world for x in rows)
        flat = [i for x in rows for i in x]
        assert flat == list(range((n // world) * world))
    full = list(SyntheticLoader(16, 2, n_classes=10, pin=False))
    parts = [list(SyntheticLoader(16, 2, n_classes=10, pin=False, rank=r, world=4)) for r in range(4)]
    for b in range(2):
        assert torch.equal(torch.cat([parts[r][b][0] for r in range(4)]), full[b][0])
        assert torch.equal(torch.cat([parts[r][b][1] for r in range(4)]), full[b][1])
    plain = list(SyntheticLoader(16, 1, pin=False, rank=1, world=2))
    assert plain[0][0].shape[0] == 8 and plain[0][1].shape[0] == 8


def test_gradient_buckets_of_the_dcgan_networks():
    """parallel.GradBuckets on the real parameter lists: reverse parameter order, >= 2 MB per bucket, and the buckets that
    complete with the last gradients merged into one collective (profiles/r02_scale_timeline.md).  D: [conv5.w .. conv4.w]
    (hidden behind the sweep) + [norm3 .. conv1.w]; G: [conv5.w .. conv3.w] + [norm2, conv2.w, norm1, conv1.w]."""
    from jck_generation_b200 import parallel
    from oracle import models
    g, d = models.build("DCGAN", seed=1)
    for net, first, tail in ((d, {"conv5.weight", "norm4.weight", "norm4.bias", "conv4.weight"}, "conv1.weight"),
                             (g, {"conv5.weight", "norm4.weight", "norm4.bias", "conv4.weight", "norm3.weight", "norm3.bias",
                                  "conv3.weight"}, "conv1.weight")):
        flat = parallel.FlatParams(net)
        names = {id(p): n for n, p in net.named_parameters()}
        merged = parallel.GradBuckets(flat, parallel.LocalComm(), merge_tail=True)
        plain = parallel.GradBuckets(flat, parallel.LocalComm(), merge_tail=False)
        assert len(plain.buckets) == 3 and len(merged.buckets) == 2
        assert {names[i] for i in merged.buckets[0][2]} == first
        assert "conv2.weight" in {names[i] for i in merged.buckets[1][2]} and tail in {names[i] for i in merged.buckets[1][2]}
        # contiguous cover of the flat buffer, every parameter in exactly one bucket
        assert merged.buckets[0][0] == 0 and merged.buckets[0][1] == merged.buckets[1][0] and merged.buckets[1][1] == flat.numel
        assert sorted(i for b in merged.buckets for i in b[2]) == sorted(names)
