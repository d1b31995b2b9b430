"""The CGAN G+D train step (reference train/cgan_trainer.py:173-213) on the sm_100a kernels.

    D(real_n, y) -> BCE(.9);  fake = G(z, y);  D(fake_n.detach(), y) -> BCE(.1)                 (:179-197)
    GP on x_hat = a*real_n + (1-a)*fake_n (both detached)                                       (:200, 114-131)
    error_d = BCE_real + BCE_fake + 10*GP;  error_d.backward();  optimizer_d.step()             (:201-204)
    D(fake_n, y) with the updated D -> BCE(.9) -> backward into G;  optimizer_g.step()          (:206-213)

Unlike DCGAN, the penalty IS back-propagated, so D's update contains second-order terms.  They come from
CganDiscriminatorEngine's explicit sweep (input-gradient sweep -> its adjoint -> one ordinary backward
with injected terms; engine_cgan.py), not from a generic autograd engine.  The three D passes of the D
update share weights and run as one 3B-image forward with three BatchNorm groups and three independent
dropout masks; every pass (also the G step's) draws its own mask, as nn.Dropout does in the reference.
"""
import torch

from .. import ops
from ..engine_cgan import P_DROP
from .dcgan_step import LABEL_FAKE, LABEL_REAL, S_FAKE, S_G, S_GP, S_REAL, DCGANStep


class CGANStep(DCGANStep):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.nz = self.g.nz            # G.conv1 takes nz + n_classes inputs; the noise itself is nz wide
        # D's gradients ACCUMULATE over three sweeps here (first order, second order, penalty), so nothing is final
        # before the last one: one exchange per network after its backward (averaged: the heads divide by the local batch)
        self.ed.grad_sync = self.eg.grad_sync = None

    def draw(self, B):
        r = super().draw(B)
        masks = torch.empty(4 * B, 256, dtype=torch.float32, device=self.dev)
        ops.dropout_mask(masks, P_DROP, self.seed, 16 * self.comm.rank + 5, self.rng_counter)
        r["drop"] = [masks[i * B:(i + 1) * B] for i in range(4)]
        r["drop_abc"] = masks[:3 * B]
        return r

    def run(self, real, labels, rng=None, after_d_update=None):
        """real [B,nc,64,64] fp32, labels [B,n_classes] one-hot (int64 or fp32) on the device.
        `after_d_update` (parity tests only) is called right after optimizer_d.step()."""
        ed, eg = self.ed, self.eg
        B = real.shape[0]
        dev, dt = self.dev, self.dtype
        r = rng if rng is not None else self.draw(B)
        self.flat_d.rebind()
        self.flat_g.rebind()
        self.arena.reset()
        lay = ed.img_layout
        if labels.dtype != torch.float32:
            lab = torch.empty(B, labels.shape[1], dtype=torch.float32, device=dev)
            ops.i64_to_f32(labels.contiguous(), lab)
        else:
            lab = labels.contiguous()

        p4 = lay == ops.IMG_P4
        X = self._p4("X", 3 * B) if p4 else ops.img_alloc(3 * B, self.nc, 64, 64, dt, dev, lay)
        real_n = torch.empty(B, self.nc, 64, 64, dtype=torch.float32, device=dev)
        drawn = not torch.is_tensor(r["noise_real"])        # instance noise drawn in registers (DCGANStep.draw)
        if drawn:
            ops.prep_image_rng(real, self.seed, r["noise_real"], self.rng_counter, 0.9, 0.1, out_nhwc=X[0:B],
                               out_nchw=real_n, layout=lay)                                                       # :182
        else:
            ops.prep_image(real, out_nhwc=X[0:B], m1=r["noise_real"], a1=0.9, b1=0.1, out_nchw=real_n, layout=lay)
        # :190 -- cat([z, labels]) goes straight into conv1's operand (GeneratorEngine.forward, jck_concat_rows)
        gctx = eg.forward(r["z"].reshape(B, self.nz).contiguous(), labels=lab,
                          y5_out=self._p4("y5", B) if eg.img_layout == ops.IMG_P4 else None)
        fake_raw = torch.empty(B, self.nc, 64, 64, dtype=torch.float32, device=dev)
        fake_n = torch.empty_like(fake_raw)
        if drawn:
            ops.g_out_fwd_rng(gctx.y[5], self.seed, r["noise_fake"], self.rng_counter, 0.9, 0.1, fake_raw, fake_n,
                              X[B:2 * B], (B, self.nc, 64, 64), layout=lay)
            ops.rng_advance(self.rng_counter, (B * self.nc * 64 * 64 + 3) // 4)
        else:
            ops.g_out_fwd(gctx.y[5], r["noise_fake"], 0.9, 0.1, fake_raw, fake_n, X[B:2 * B], (B, self.nc, 64, 64), layout=lay)
        ops.prep_image(real_n, out_nhwc=X[2 * B:3 * B], a1=1.0, x2=fake_n, alpha=r["alpha"].reshape(B), layout=lay)  # :115

        scal = self.arena.take(8).view(4, 2)               # valid until the next step's reset (callers clone to keep it)
        masks = r.get("drop_abc")
        if masks is None:                                  # injected masks: three separate [B,256] blocks
            masks = torch.cat([m.reshape(B, 256).float() for m in r["drop"][:3]]).contiguous()
        ctx = ed.trunk_forward(X, groups=3)                                                                       # :184,194,118
        ed.head_forward(ctx, lab, masks, targets=[LABEL_REAL, LABEL_FAKE, None], scalars=scal)

        # ---- D update: first + second order ------------------------------------------------------------------
        ops.zero(self.flat_d.grad)                         # everything below accumulates
        cc = ctx.slice(2, 3)
        cc.head = {k: (v[2 * B:3 * B] if (torch.is_tensor(v) and v.shape[0] == 3 * B) else v) for k, v in ctx.head.items()}
        g_a4 = ed.head_gp_seed(cc)
        v = ed.trunk_backward(cc, g_a4, wgrad=False, input_grad=True, fuse=False,                                 # :120-127
                              dx_out=self._p4("dx", B) if p4 else None)
        u = self._p4("u", B) if p4 else torch.empty_like(v)
        world_b = B * self.comm.world_size
        ops.gp_seed(v, u, scal[S_GP], self.lambda_gp * 2.0 / B)                                                # :130, 201
        sbar, ybar = ed.adjoint_sweep(cc, u)
        dls = torch.empty(3 * B, dtype=torch.float32, device=dev)
        ops.logit_grad(ctx.prob[0:B], dls[0:B], mode=0, target=LABEL_REAL, scale=1.0 / B)
        ops.logit_grad(ctx.prob[B:2 * B], dls[B:2 * B], mode=0, target=LABEL_FAKE, scale=1.0 / B)
        ops.copy_f32(sbar.contiguous(), dls[2 * B:3 * B])
        da4 = ed.head_backward(ctx, dls, wgrad=True)
        ed.flush_linear1_grad(accumulate=True)
        ed.trunk_backward(ctx, da4, wgrad=True, input_grad=False, accumulate=True, inject=ybar, inject_rows=(2 * B, 3 * B))  # :203
        ed.join_wgrad()
        self.comm.allreduce_mean_(self.flat_d.grad)
        self.opt_d.step()                                                                                         # :204
        if after_d_update is not None:
            after_d_update()
        ed.refresh(force=True)

        # ---- G update ---------------------------------------------------------------------------------------------
        ctx2 = ed.trunk_forward(X[B:2 * B], groups=1)                                                             # :209
        ed.head_forward(ctx2, lab, r["drop"][3].reshape(B, 256).contiguous().float(), targets=[LABEL_REAL], scalars=scal[S_G:S_G + 1])
        dls2 = torch.empty(B, dtype=torch.float32, device=dev)
        ops.logit_grad(ctx2.prob, dls2, mode=0, target=LABEL_REAL, scale=1.0 / B)
        da4 = ed.head_backward(ctx2, dls2, wgrad=False)
        dmix = ed.trunk_backward(ctx2, da4, wgrad=False, input_grad=True, dx_out=self._p4("dmix", B) if p4 else None)  # :211
        dy5 = self._p4("dy5", B) if p4 else torch.empty_like(dmix)
        ops.g_out_bwd(dmix, fake_raw, 0.9, dy5, layout=lay)
        eg.backward(gctx, dy5, accumulate=False)
        eg.join_wgrad()
        self.comm.allreduce_mean_(self.flat_g.grad)
        self.opt_g.step()                                                                                         # :213
        eg.refresh(force=True)
        self.last = {"fake_raw": fake_raw, "gp_grad_nhwc": v, "ctx": ctx, "ctx_g": gctx, "ctx_d": ctx2, "world_b": world_b}
        return scal

    # ---- CUDA graph: DCGANStep.capture() with (real, labels) as the static inputs --------------------------------
    def _make_static(self, batch):
        super()._make_static(batch)
        self._static_labels = torch.zeros(batch, self.g.n_classes, dtype=self._label_dtype, device=self.dev)
        self._static_labels[:, 0] = 1                       # a valid one-hot for the warm-up steps

    def _run_static(self):
        return self.run(self._static, self._static_labels)

    def capture(self, batch, label_dtype=torch.int64):
        self._label_dtype = label_dtype
        return super().capture(batch)

    def replay(self, real, labels):
        self._static_labels.copy_(labels, non_blocking=True)
        return super().replay(real)
