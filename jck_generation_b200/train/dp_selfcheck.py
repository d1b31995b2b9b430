"""Data-parallel self-check: W ranks, each on B/W rows of one global batch (SyncBN statistics over the global batch,
summed gradient buckets), must reproduce ONE rank's step on all B rows -- which tests/test_gpu_step.py in turn pins to the
CPU oracle.  Run by `bench.py` (outside the timed region, every N > 1) and by tests/dp_check.py.

The comparison runs at lr = 0: Adam's first update is lr*g/(|g|+eps), i.e. sign-like, so gradient elements below fp32
summation-order noise flip by 2*lr between ANY two runs and everything computed after optimizer_d.step() in the same step
would measure that coin toss instead of the collectives.  With lr = 0 the whole step (four D passes, G, all gradients)
is a pure function of the inputs and must agree: forward activations / BatchNorm statistics / losses tightly, D's gradients
within the LeakyReLU-kink noise, G's gradients within the BatchNorm-backward amplification of it."""
import torch

from .. import parallel
from ..model import DCGAN
from .dcgan_step import DCGANStep
from .optim import FusedAdam

TOL = {  # dtype -> (forward / statistics, D gradients, pass-D-derived gradients)
    # fp32 D gradients are bimodal between ANY two runs of the same binary: ~2e-6 when no LeakyReLU pre-activation sits within
    # summation-order noise of zero, 4e-4 .. 2e-3 when one or two do and take the other branch on one side (the per-channel
    # statistics are summed in a different order across ranks; DESIGN.md section 2, "kink flips")
    torch.float32: (1e-4, 5e-3, 2e-2),
    torch.bfloat16: (1e-1, 1e-1, 2e-1),
}


def _build(dtype, comm, seed):
    torch.manual_seed(seed)                     # same seed on every rank -> identical replicas
    g, d = DCGAN.Generator(dtype=dtype), DCGAN.Discriminator(dtype=dtype)
    g.apply(DCGAN.weights_init)
    d.apply(DCGAN.weights_init)
    g = g.cuda().set_compute(dtype=dtype, comm=comm)
    d = d.cuda().set_compute(dtype=dtype, comm=comm)
    fg, fd = parallel.FlatParams(g), parallel.FlatParams(d)
    og = FusedAdam(g.parameters(), lr=0.0, betas=[0.5, 0.999], flat=fg)
    od = FusedAdam(d.parameters(), lr=0.0, betas=[0.5, 0.999], flat=fd)
    return g, d, DCGANStep(g, d, og, od, fg, fd, comm)


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def run(comm, per_rank=32, dtypes=(torch.float32, torch.bfloat16), seed=12345):
    """Returns {"ok": bool, "world": W, "global_batch": B, "<dtype>": {"forward": e, "d_grads": e, "g_grads": e, ...}} on
    rank 0 (None elsewhere).  Every rank must call it."""
    W = comm.world_size
    B = per_rank * W
    gen = torch.Generator().manual_seed(seed + 1)
    real = (torch.rand(B, 3, 64, 64, generator=gen) * 2 - 1).cuda()
    rng = {"noise_real": torch.randn(B, 3, 64, 64, generator=gen).cuda(), "z": torch.randn(B, 100, 1, 1, generator=gen).cuda(),
           "noise_fake": torch.randn(B, 3, 64, 64, generator=gen).cuda(), "alpha": torch.rand(B, 1, 1, 1, generator=gen).cuda()}
    out = {"ok": True, "world": W, "global_batch": B, "syncbn_transport": comm.transport}
    for dtype in dtypes:
        g, d, step = _build(dtype, comm, seed)
        shard = {k: parallel.shard_rows(v, comm).contiguous() for k, v in rng.items()}
        scal = step.run(parallel.shard_rows(real, comm).contiguous(), shard).clone()
        comm.allreduce_mean_(scal)
        torch.cuda.synchronize()
        if comm.rank == 0:
            g1, d1, step1 = _build(dtype, parallel.LocalComm(), seed)
            scal1 = step1.run(real, rng)
            torch.cuda.synchronize()
            fwd = {"scalars": _rel(scal, scal1)}
            for tag, m, m1 in (("d.", d, d1), ("g.", g, g1)):
                for (n, p), (_, q) in zip(m.state_dict().items(), m1.state_dict().items()):
                    if n.endswith("num_batches_tracked"):
                        fwd[tag + n] = float(abs(int(p) - int(q)))
                    elif "running" in n:
                        fwd[tag + n] = _rel(p, q)
            for k in (1, 2, 3, 4):
                fwd[f"G.y{k}"] = _rel(step.last["ctx_g"].y[k], step1.last["ctx_g"].y[k][:per_rank])
                fwd[f"passD.y{k}"] = _rel(step.last["ctx_d"].y[k], step1.last["ctx_d"].y[k][:per_rank])
            dg = {n: _rel(p.grad, q.grad) for (n, p), (_, q) in zip(d.named_parameters(), d1.named_parameters())}
            gg = {n: _rel(p.grad, q.grad) for (n, p), (_, q) in zip(g.named_parameters(), g1.named_parameters())}
            tf, td, tg = TOL[dtype]
            res = {"forward": max(fwd.values()), "d_grads": max(dg.values()), "g_grads": max(gg.values()),
                   "tol": [tf, td, tg]}
            res["ok"] = bool(res["forward"] <= tf and res["d_grads"] <= td and res["g_grads"] <= tg)
            out[str(dtype).replace("torch.", "")] = res
            out["ok"] = out["ok"] and res["ok"]
        comm.barrier()
        del g, d, step
        torch.cuda.empty_cache()
    return out if comm.rank == 0 else None
