"""The DCGAN G+D train step as one fixed launch sequence of sm_100a kernels.

Semantics: exactly the body of the reference's loop, train/dcgan_trainer.py:155-189 --
  A  D(0.9*real + 0.1*n1) vs label .9  -> backward          (:155-165)
  B  fake = G(z); D((0.9*fake + 0.1*n2).detach()) vs .1 -> backward, grads accumulate on A   (:168-176)
  C  gradient penalty on x_hat = a*real_n + (1-a)*fake_n: a D forward in train mode (it moves the BN
     running statistics) and an input-gradient sweep; the value is only logged (:178-179)
     optimizer_d.step()                                                                      (:180)
  D  D(fake_n) with the updated D vs label .9 -> backward into G; optimizer_g.step()       (:182-189)
What differs is scheduling, not arithmetic: passes A, B, C read the same D weights, so they run as ONE
3B-image forward with three BatchNorm statistic groups (running statistics updated in A, B, C order);
the reference's wasted D weight-gradient of pass D (zeroed at :155 before any use) is not computed.

Random tensors: `rng` injects host-made tensors (parity mode -- CPU mt19937 and Philox streams can
never agree); otherwise they are drawn on the device by our Philox kernel.
"""
import os

import torch

from .. import ops
from ..parallel import AuxComm, GradBuckets

LABEL_REAL = 0.9
LABEL_FAKE = 0.1

# rows of the per-step scalar block
S_REAL, S_FAKE, S_GP, S_G = 0, 1, 2, 3


def _null():
    import contextlib
    return contextlib.nullcontext()


class DCGANStep:
    def __init__(self, model_g, model_d, opt_g, opt_d, flat_g, flat_d, comm, lambda_gp=10.0, seed=12345):
        self.g, self.d = model_g, model_d
        self.eg, self.ed = model_g.engine(), model_d.engine()
        self.opt_g, self.opt_d = opt_g, opt_d
        self.flat_g, self.flat_d = flat_g, flat_d
        self.comm = comm
        self.lambda_gp = lambda_gp
        self.seed = seed
        self.dev = self.ed.dev
        self.dtype = self.ed.dtype
        self.nc = self.ed.nc
        self.nz = self.eg.K1
        self.rng_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._graph = None
        self._static = None
        self._p4_bufs = {}
        # every fp32 accumulator of a step lives in one arena zeroed by one memset (no torch fill kernels in the step)
        self.arena = ops.ZeroArena(self.dev)
        self.eg.arena = self.ed.arena = self.arena
        # weight-gradient kernels run beside the sweep that produces their operands (engine._GradTarget)
        self.gp_stream = None
        if self.dtype == torch.bfloat16:
            self.eg.wgrad_stream = self.ed.wgrad_stream = torch.cuda.Stream(device=self.dev)
            # the gradient-penalty sweep (pass C) is independent of the A+B backward: it runs on its own stream, its
            # HBM-bound BatchNorm passes beside the other sweep's convolutions and vice versa (JCK_GP_STREAM=0: in line)
            if os.environ.get("JCK_GP_STREAM", "1") != "0":
                self.gp_stream = torch.cuda.Stream(device=self.dev)
        self.comm_gp = AuxComm(comm) if (self.gp_stream is not None and comm.world_size > 1) else None
        # gradient exchange: per-bucket all-reduce started as the bucket's last gradient is written (parallel.GradBuckets)
        self.sync_d = GradBuckets(flat_d, comm)
        self.sync_g = GradBuckets(flat_g, comm)
        if comm.world_size > 1:
            self.ed.grad_sync, self.eg.grad_sync = self.sync_d, self.sync_g

    # ---- random tensors ----------------------------------------------------------------------------
    def draw(self, B):
        """Device-side draws.  z and alpha are materialised; the two instance-noise tensors (the bulk: 2 x B*nc*64*64
        Gaussians) are NOT -- they are named by their Philox stream id and drawn in registers by the kernels that
        consume them (ops.prep_image_rng / ops.g_out_fwd_rng produce exactly what ops.randn would have stored).
        run() advances the counter once every consumer of this step's streams has been queued."""
        sid = 16 * self.comm.rank
        dev = self.dev
        r = {"noise_real": sid + 1,
             "z": torch.empty(B, self.nz, 1, 1, device=dev),
             "noise_fake": sid + 3,
             "alpha": torch.empty(B, 1, 1, 1, device=dev)}
        ops.randn(r["z"], self.seed, sid + 2, self.rng_counter)
        ops.rand(r["alpha"], self.seed, sid + 4, self.rng_counter)
        return r

    def _p4(self, tag, B):
        """Image-side buffers in the zero-bordered JCK_IMG_P4 layout, kept across steps: every kernel writes interior
        pixels only (pad channel = 0), so the border is zeroed once instead of once per step."""
        key = (tag, B)
        buf = self._p4_bufs.get(key)
        if buf is None:
            buf = self._p4_bufs[key] = ops.img_alloc(B, self.nc, 64, 64, self.dtype, self.dev, ops.IMG_P4)
        return buf

    # ---- the step ------------------------------------------------------------------------------------
    def run(self, real, rng=None, after_d_update=None):
        """real: [B,nc,64,64] fp32 on the device (this rank's rows).  Returns a [4,2] fp32 device tensor:
        rows (real, fake, gp, g), columns (BCE mean | GP value, mean D output) -- local-batch means.
        `after_d_update` (parity tests only) is called right after optimizer_d.step()."""
        ed, eg = self.ed, self.eg
        B = real.shape[0]
        dev, dt = self.dev, self.dtype
        r = rng if rng is not None else self.draw(B)
        self.flat_d.rebind()
        self.flat_g.rebind()
        self.arena.reset()

        lay = ed.img_layout
        p4 = lay == ops.IMG_P4
        X = self._p4("X", 3 * B) if p4 else ops.img_alloc(3 * B, self.nc, 64, 64, dt, dev, lay)   # [real_n | fake_n | x_hat]
        real_n = torch.empty(B, self.nc, 64, 64, dtype=torch.float32, device=dev)
        drawn = not torch.is_tensor(r["noise_real"])
        if drawn:
            ops.prep_image_rng(real, self.seed, r["noise_real"], self.rng_counter, 0.9, 0.1, out_nhwc=X[0:B],
                               out_nchw=real_n, layout=lay)                                                # :160
        else:
            ops.prep_image(real, out_nhwc=X[0:B], m1=r["noise_real"], a1=0.9, b1=0.1, out_nchw=real_n, layout=lay)

        gctx = eg.forward(r["z"].reshape(B, self.nz), y5_out=self._p4("y5", B) if eg.img_layout == ops.IMG_P4 else None)  # :169
        fake_raw = torch.empty(B, self.nc, 64, 64, dtype=torch.float32, device=dev)
        fake_n = torch.empty_like(fake_raw)
        if drawn:
            ops.g_out_fwd_rng(gctx.y[5], self.seed, r["noise_fake"], self.rng_counter, 0.9, 0.1, fake_raw, fake_n,
                              X[B:2 * B], (B, self.nc, 64, 64), layout=lay)                                # :171
            ops.rng_advance(self.rng_counter, (B * self.nc * 64 * 64 + 3) // 4)
        else:
            ops.g_out_fwd(gctx.y[5], r["noise_fake"], 0.9, 0.1, fake_raw, fake_n, X[B:2 * B],
                          (B, self.nc, 64, 64), layout=lay)
        ops.prep_image(real_n, out_nhwc=X[2 * B:3 * B], a1=1.0, x2=fake_n, alpha=r["alpha"].reshape(B),
                       layout=lay)                                                                         # :112

        scal = self.arena.take(8).view(4, 2)               # valid until the next step's reset (callers clone to keep it)
        ctx = ed.trunk_forward(X, groups=3)                                                                # :162,173,114
        ed.head_forward(ctx, targets=[LABEL_REAL, LABEL_FAKE, None], scalars=scal)

        # pass C, the penalty's input-gradient sweep (:116-126): forked onto its own stream
        main = torch.cuda.current_stream()
        gps = self.gp_stream
        if gps is not None:
            gps.wait_stream(main)
        with torch.cuda.stream(gps) if gps is not None else _null():
            cc = ctx.slice(2, 3)
            da4c = ed.head_backward(cc, mode=1, wgrad=False)
            dx = ed.trunk_backward(cc, da4c, wgrad=False, input_grad=True, dx_out=self._p4("dx", B) if p4 else None,
                                   comm=self.comm_gp)
            ops.gp_penalty(dx, scal[S_GP])

        # passes A+B (:164,175): D's parameter gradients; every bucket of them is exchanged as soon as it is final,
        # while the rest of this sweep and the penalty sweep run
        self.sync_d.begin()
        cab = ctx.slice(0, 2)
        da4 = ed.head_backward(cab, mode=0, targets=[LABEL_REAL, LABEL_FAKE], wgrad=True, accumulate=False)
        ed.trunk_backward(cab, da4, wgrad=True, input_grad=False, accumulate=False)
        ed.join_wgrad()
        self.sync_d.finish()
        self.opt_d.step()                                                                                  # :180
        if after_d_update is not None:
            after_d_update()
        if gps is not None:
            main.wait_stream(gps)       # the penalty sweep reads the packed weights refresh() is about to overwrite
        ed.refresh(force=True)

        ctx2 = ed.trunk_forward(X[B:2 * B], groups=1)                                                      # :185
        ed.head_forward(ctx2, targets=[LABEL_REAL], scalars=scal[S_G:S_G + 1])
        da4 = ed.head_backward(ctx2, mode=0, targets=[LABEL_REAL], wgrad=False)                            # :187
        dmix = ed.trunk_backward(ctx2, da4, wgrad=False, input_grad=True, dx_out=self._p4("dmix", B) if p4 else None)
        dy5 = self._p4("dy5", B) if p4 else torch.empty_like(dmix)
        ops.g_out_bwd(dmix, fake_raw, 0.9, dy5, layout=lay)
        self.sync_g.begin()
        eg.backward(gctx, dy5, accumulate=False)
        eg.join_wgrad()
        self.sync_g.finish()
        self.opt_g.step()                                                                                  # :189
        eg.refresh(force=True)
        self.last = {"fake_raw": fake_raw, "gp_grad_nhwc": dx, "ctx": ctx, "ctx_g": gctx, "ctx_d": ctx2,
                     "dmix": dmix, "dy5": dy5}
        return scal

    # ---- CUDA graph ---------------------------------------------------------------------------------
    def capture(self, batch):
        """Capture one step (device-drawn random tensors) into a CUDA graph; replay() then costs one
        launch.  Under data parallelism the NCCL all-reduces (SyncBN statistics, gradient buckets) are
        captured with it: every rank replays the same sequence, so the collectives still pair up."""
        self._make_static(batch)
        # warm-up steps really train; put the training state back afterwards
        bufs = [b for m in (self.g, self.d) for b in m.buffers()]
        keep = [t.clone() for t in (self.flat_g.flat, self.flat_g.exp_avg, self.flat_g.exp_avg_sq,
                                    self.flat_d.flat, self.flat_d.exp_avg, self.flat_d.exp_avg_sq,
                                    self.opt_g.step_dev, self.opt_d.step_dev, self.rng_counter, *bufs)]
        steps = (self.opt_g.steps_done, self.opt_d.steps_done)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):                      # warm-up: allocator pools, smem attributes, packs
                self._run_static()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._graph_out = self._run_static()
        for dst, src in zip((self.flat_g.flat, self.flat_g.exp_avg, self.flat_g.exp_avg_sq,
                             self.flat_d.flat, self.flat_d.exp_avg, self.flat_d.exp_avg_sq,
                             self.opt_g.step_dev, self.opt_d.step_dev, self.rng_counter, *bufs), keep):
            dst.copy_(src)
        self.opt_g.steps_done, self.opt_d.steps_done = steps
        self.eg.refresh(force=True)
        self.ed.refresh(force=True)
        torch.cuda.synchronize()
        return self

    # static inputs of the captured step (subclasses with more inputs override these two)
    def _make_static(self, batch):
        self._static = torch.zeros(batch, self.nc, 64, 64, dtype=torch.float32, device=self.dev)

    def _run_static(self):
        return self.run(self._static)

    def replay(self, real):
        self._static.copy_(real, non_blocking=True)
        self._graph.replay()
        self.opt_d.steps_done += 1
        self.opt_g.steps_done += 1
        return self._graph_out

    @staticmethod
    def summarize(scal, lambda_gp=10.0):
        """host view of a scalar block: the quantities the reference logs (:191-196)."""
        s = scal.detach().float().cpu()
        return {"loss_d": float(s[S_REAL, 0] + s[S_FAKE, 0] + lambda_gp * s[S_GP, 0]),
                "loss_g": float(s[S_G, 0]), "x_d": float(s[S_REAL, 1]), "z1_gd": float(s[S_FAKE, 1]),
                "z2_gd": float(s[S_G, 1]), "gp": float(s[S_GP, 0]),
                "err_real": float(s[S_REAL, 0]), "err_fake": float(s[S_FAKE, 0])}
