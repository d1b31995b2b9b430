"""Data-parallel equivalence on real GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 -m tests.dp_check

W ranks, each on B/W rows of the same global batch with SyncBN + averaged gradients, must reproduce the
single-GPU step on all B rows (fp32 arithmetic: tight; bf16: loose), which tests/test_gpu_step.py in
turn pins to the CPU oracle."""
import sys

import torch

from jck_generation_b200 import parallel
from jck_generation_b200.model import DCGAN
from jck_generation_b200.train.dcgan_step import DCGANStep
from jck_generation_b200.train.optim import FusedAdam
from oracle import models as omodels
from oracle import steps as osteps


def build(dtype, comm, lr=2e-4):
    g_o, d_o = omodels.build("DCGAN", seed=12345)
    g = DCGAN.Generator(dtype=dtype).cuda().set_compute(dtype=dtype, comm=comm)
    d = DCGAN.Discriminator(dtype=dtype).cuda().set_compute(dtype=dtype, comm=comm)
    g.load_state_dict(g_o.state_dict()); d.load_state_dict(d_o.state_dict())
    fg, fd = parallel.FlatParams(g), parallel.FlatParams(d)
    og = FusedAdam(g.parameters(), lr=lr, betas=[0.5, 0.999], flat=fg)
    od = FusedAdam(d.parameters(), lr=lr, betas=[0.5, 0.999], flat=fd)
    return g, d, DCGANStep(g, d, og, od, fg, fd, comm)


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    comm = parallel.init_from_env()
    B = 16 * comm.world_size
    real = osteps.make_real(B, n_steps=1)[0].cuda()
    rng = {k: v.cuda() for k, v in osteps.make_rng(B, n_steps=1, seed=5)[0].items()}
    worst = {}
    for dtype, tol in ((torch.float32, 2e-4), (torch.bfloat16, 6e-2)):
        g, d, step = build(dtype, comm)
        shard = {k: parallel.shard_rows(v, comm).contiguous() for k, v in rng.items()}
        scal = step.run(parallel.shard_rows(real, comm).contiguous(), shard).clone()
        comm.allreduce_mean_(scal)
        torch.cuda.synchronize()
        if comm.rank == 0:
            g1, d1, step1 = build(dtype, parallel.LocalComm())
            scal1 = step1.run(real, rng)
            torch.cuda.synchronize()
            errs = {"scalars": rel(scal, scal1)}
            for (n, p), (_, q) in zip(list(d.state_dict().items()) + list(g.state_dict().items()),
                                      list(d1.state_dict().items()) + list(g1.state_dict().items())):
                if n.endswith("num_batches_tracked"):
                    assert int(p) == int(q), n
                elif "running" in n or dtype == torch.float32:
                    errs[n] = rel(p, q)
            w = max(errs.items(), key=lambda kv: kv[1])
            worst[str(dtype)] = w
            print(f"dp_check {dtype} world={comm.world_size}: worst {w[0]} = {w[1]:.3e} (tol {tol})", flush=True)
            assert w[1] <= tol, errs
        comm.barrier()
    if comm.rank == 0:
        print("dp_check OK", worst, flush=True)
    if comm.world_size > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
