"""CGAN discriminator on the sm_100a kernels: the DCGAN trunk (engine.DiscriminatorEngine) plus the
conditional head of model/CGAN.py:79-123

    lab = LeakyReLU(Linear(n_classes -> E)(labels.float()))                       (:83-84, 112)
    out = Sigmoid(Linear(256 -> 1)(Dropout(.25)(Linear(16*C4 + E -> 256)(cat[flatten(a4), lab]))))   (:118-122)

and the second-order sweep the CGAN discriminator update needs, because its gradient penalty is
back-propagated (train/cgan_trainer.py:200-204).  The sweep is explicit (no generic double-backward
engine); tests/notes/cgan_second_order_check.py proves the algorithm against torch.autograd on the CPU:

  1. input-gradient sweep  v = d sum(D(x_hat)) / d x_hat   (head_gp_seed + trunk_backward, intermediates kept)
  2. u = d(lambda*GP)/dv, then the ADJOINT of sweep 1 from the image side up to the head (adjoint_sweep):
     per layer  dbar_k = conv_k(abar_{k-1}),  dW_k += wgrad(d_k, abar_{k-1}),  BatchNorm-backward adjoint
     -> (abar_k, ybar_k, dgamma_k);  head adjoint -> dW1, dw2 and the logit term sbar = gbar_s * sigma''(s)
  3. ONE ordinary backward over all three D passes with per-row logit gradients [BCE real | BCE fake | sbar]
     and ybar_k injected at the raw conv outputs of the penalty rows.

The three large products of the head -- with the 16*C4 = 8192-wide feature vector: linear1 forward (x.W^T), its
input gradient (g.W) and weight gradient (g^T.x), and their twins in the second-order sweep -- run on tcgen05
(jck_gemm_tc: K-major and MN-major operands, so no transposed copies; split-K where the contraction is long) in bf16
mode and on the exact CUDA-core jck_dense kernel in fp32 parity mode; everything else is small fp32 row ops."""
import torch

from . import ops
from .engine import DiscriminatorEngine, LRELU, _Workspace

P_DROP = 0.25
KEEP_SCALE = 1.0 / (1.0 - P_DROP)


class CganDiscriminatorEngine(DiscriminatorEngine):
    def __init__(self, module, dtype=torch.bfloat16, comm=None, algo=ops.ALGO_AUTO):
        super().__init__(module, dtype, comm, algo)
        dev = self.dev
        self.C4 = self.convs[4].Ca
        self.F = 16 * self.C4                                   # flattened trunk features
        self.H = int(module.linear1.weight.shape[0])            # 256
        self.E = int(module.label_embedding.weight.shape[0])    # 200
        self.ncls = int(module.label_embedding.weight.shape[1])
        self.w1a = torch.empty(self.H, self.F, dtype=dtype, device=dev)       # NHWC column order
        self.w1b = torch.empty(self.H, self.E, dtype=torch.float32, device=dev)
        self.dw1a = torch.zeros(self.H, self.F, dtype=torch.float32, device=dev)
        self.dw1b = torch.zeros(self.H, self.E, dtype=torch.float32, device=dev)
        self.ones = torch.ones(1 << 16, dtype=torch.float32, device=dev)
        self._w1_seen = None
        self.tc_head = dtype == torch.bfloat16 and algo != ops.ALGO_SIMT
        self.gws = _Workspace(dev)          # split-K partials of the head GEMMs (self.ws belongs to the conv wgrads)

    # parameters (detached views)
    def _p(self, name):
        mod, attr = name.split(".")
        return getattr(getattr(self.m, mod), attr).detach()

    def refresh(self, force=False):
        super().refresh(force)
        w = self.m.linear1.weight
        key = (w._version, w.data_ptr())
        if force or key != self._w1_seen:
            ops.pack_linear(w.detach(), self.w1a, self.w1b, self.C4, 16)
            self._w1_seen = key

    # ---- the head's large products ----------------------------------------------------------------------------
    def _bf(self, x):
        out = torch.empty(x.shape, dtype=torch.bfloat16, device=self.dev)
        ops.f32_to_bf16(x, out)
        return out

    def _gemm(self, A, a_mn, lda, Bm, b_mn, ldb, C, M, N, K, accumulate=False):
        nbytes = ops.gemm_tc_workspace_bytes(M, N, K)
        ops.gemm_tc(A, a_mn, lda, Bm, b_mn, ldb, C, M, N, K, accumulate=accumulate,
                    workspace=self.gws.get(nbytes) if nbytes else None)

    def _feat_fwd(self, x, out, B):
        """out[B,H] = x[B,F] . w1a^T"""
        if self.tc_head:
            self._gemm(x, 0, self.F, self.w1a, 0, self.F, out, B, self.H, self.F)
        else:
            ops.dense(x, self.F, 1, self.w1a, self.F, 1, out, B, self.H, self.F)

    def _feat_dgrad(self, g_h, out, B):
        """out[B,F] = g_h[B,H] . w1a   (g_h fp32; rounded to bf16 for the tensor cores like every other gradient)"""
        if self.tc_head:
            self._gemm(self._bf(g_h), 0, self.H, self.w1a, 1, self.F, out, B, self.F, self.H)
        else:
            ops.dense(g_h, self.H, 1, self.w1a, 1, self.F, out, B, self.F, self.H)

    def _feat_wgrad(self, g_h, x, B):
        """dw1a[H,F] += g_h^T[H,B] . x[B,F]"""
        if self.tc_head:
            self._gemm(self._bf(g_h), 1, self.H, x, 1, self.F, self.dw1a, self.H, self.F, B, accumulate=True)
        else:
            ops.dense(g_h, 1, self.H, x, 1, self.F, self.dw1a, self.H, self.F, B, accumulate=True)

    # ---- small helpers -------------------------------------------------------------------------------------
    def _colsum_into(self, x, out, rows, cols):
        """out[cols] += sum over rows of x[rows][cols]  (a [1 x rows] . [rows x cols] product with ones)"""
        ops.dense(self.ones, 0, 1, x, 1, cols, out, 1, cols, rows, accumulate=True)

    def _f32(self, *shape):
        return torch.empty(*shape, dtype=torch.float32, device=self.dev)

    # ---- forward --------------------------------------------------------------------------------------------
    def head_forward(self, ctx, labels, masks, targets=None, scalars=None):
        """labels: [Bg, n_classes] one-hot (int64 as the reference's loader yields, or float); the same rows are
        used by every group.  masks: [B, 256] fp32 dropout keep-masks (one independent block per group)."""
        self.refresh()
        B, G = ctx.B, ctx.groups
        Bg = B // G
        hd = ctx.head
        if labels.dtype != torch.float32:
            lab = self._f32(Bg, self.ncls)
            ops.i64_to_f32(labels.contiguous(), lab)
        else:
            lab = labels.contiguous()
        We, be = self._p("label_embedding.weight"), self._p("label_embedding.bias")
        e_lin = self._f32(Bg, self.E)
        ops.dense(lab, self.ncls, 1, We, self.ncls, 1, e_lin, Bg, self.E, self.ncls)
        ops.rowop(ops.ROW_BIAS_ACT, e_lin, be, e_lin, Bg, self.E, s=1.0)
        e = self._f32(Bg, self.E)
        ops.rowop(ops.ROW_BIAS_ACT, e_lin, None, e, Bg, self.E, s=LRELU)
        he = self._f32(Bg, self.H)
        ops.dense(e, self.E, 1, self.w1b, self.E, 1, he, Bg, self.H, self.E)
        ops.rowop(ops.ROW_BIAS_ACT, he, self._p("linear1.bias"), he, Bg, self.H, s=1.0)
        a4 = ctx.a[4].view(B, self.F)
        h = self._f32(B, self.H)
        self._feat_fwd(a4, h, B)
        ops.rowop(ops.ROW_ADD_BCAST, h, he, h, B, self.H, rows_y=Bg)
        hdrop = self._f32(B, self.H)
        ops.rowop(ops.ROW_MUL, h, masks, hdrop, B, self.H, s=KEEP_SCALE)
        w2, b2 = self._p("linear2.weight"), self._p("linear2.bias")
        logit = self._f32(B, 1)
        ops.dense(hdrop, self.H, 1, w2, self.H, 1, logit, B, 1, self.H)
        ops.rowop(ops.ROW_BIAS_ACT, logit, b2, logit, B, 1, s=1.0)
        prob = self._f32(B)
        for g in range(G):
            sl = slice(g * Bg, (g + 1) * Bg)
            t = targets[g] if (targets is not None and targets[g] is not None) else 0.0
            sc = scalars[g] if (scalars is not None and targets is not None and targets[g] is not None) else None
            ops.sigmoid_bce(logit[sl].view(-1), prob[sl], t, sc)
        hd.update(lab=lab, e_lin=e_lin, e=e, hdrop=hdrop, masks=masks)
        ctx.prob = prob
        return prob

    # ---- first-order backward of the head ---------------------------------------------------------------------
    def head_backward(self, ctx, dls, wgrad=True):
        """dls: [B] fp32 = d(loss)/d(logit) per row.  Returns d/d(a4) (activation dtype); with `wgrad`
        ACCUMULATES the gradients of linear1, linear2 and label_embedding (callers zero the buffers)."""
        B, G = ctx.B, ctx.groups
        Bg = B // G
        hd = ctx.head
        w2 = self._p("linear2.weight")
        g_hd = self._f32(B, self.H)
        ops.rowop(ops.ROW_OUTER, dls, w2.view(-1), g_hd, B, self.H)
        g_h = self._f32(B, self.H)
        ops.rowop(ops.ROW_MUL, g_hd, hd["masks"], g_h, B, self.H, s=KEEP_SCALE)
        da4 = torch.empty(B, self.F, dtype=self.dtype, device=self.dev)
        self._feat_dgrad(g_h, da4, B)
        if wgrad:
            ops.dense(dls, 0, 1, hd["hdrop"], 1, self.H, self._gb(self.m.linear2.weight), 1, self.H, B, accumulate=True)
            self._colsum_into(dls.view(B, 1), self._gb(self.m.linear2.bias), B, 1)
            self._feat_wgrad(g_h, ctx.a[4].view(B, self.F), B)
            g_he = self._f32(Bg, self.H)
            ops.rowop(ops.ROW_SUM_GROUPS, g_h, g_h, g_he, B, self.H, rows_y=Bg)
            ops.dense(g_he, 1, self.H, hd["e"], 1, self.E, self.dw1b, self.H, self.E, Bg, accumulate=True)
            self._colsum_into(g_he, self._gb(self.m.linear1.bias), Bg, self.H)
            g_e = self._f32(Bg, self.E)
            ops.dense(g_he, self.H, 1, self.w1b, 1, self.E, g_e, Bg, self.E, self.H)
            ops.rowop(ops.ROW_ACT_BWD, g_e, hd["e_lin"], g_e, Bg, self.E, s=LRELU)
            ops.dense(g_e, 1, self.E, hd["lab"], 1, self.ncls, self._gb(self.m.label_embedding.weight), self.E, self.ncls, Bg,
                      accumulate=True)
            self._colsum_into(g_e, self._gb(self.m.label_embedding.bias), Bg, self.E)
        return da4.view(B, 4, 4, self.C4)

    def flush_linear1_grad(self, accumulate):
        """dw1a / dw1b (kernel layout) -> linear1.weight.grad (reference layout), then clear them."""
        ops.unpack_linear_grad(self.dw1a, self.dw1b, self._gb(self.m.linear1.weight), self.C4, 16, accumulate)
        ops.zero(self.dw1a)
        ops.zero(self.dw1b)

    # ---- gradient-penalty sweeps ---------------------------------------------------------------------------------
    def head_gp_seed(self, ctx):
        """upstream ones on the sigmoid output: g_s = p(1-p), g_h, and d sum(p)/d(a4)."""
        B = ctx.B
        hd = ctx.head
        g_s = self._f32(B)
        ops.logit_grad(ctx.prob, g_s, mode=1)
        g_hd = self._f32(B, self.H)
        ops.rowop(ops.ROW_OUTER, g_s, self._p("linear2.weight").view(-1), g_hd, B, self.H)
        g_h = self._f32(B, self.H)
        ops.rowop(ops.ROW_MUL, g_hd, hd["masks"], g_h, B, self.H, s=KEEP_SCALE)
        g_a4 = torch.empty(B, self.F, dtype=self.dtype, device=self.dev)
        self._feat_dgrad(g_h, g_a4, B)
        hd.update(g_s=g_s, g_h=g_h)
        return g_a4.view(B, 4, 4, self.C4)

    def adjoint_sweep(self, ctx, u):
        """Adjoint of the input-gradient sweep recorded in `ctx` (one statistics group), seeded with
        u = d(lambda*GP)/dv at the image.  ACCUMULATES the second-order parameter gradients (conv weights,
        BatchNorm gamma, linear1, linear2) and returns (sbar [B] = d/d(logit), {k: ybar_k})."""
        assert ctx.groups == 1
        B = ctx.B
        world = self.comm.world_size
        abar, ybar = u, {}
        for k in range(1, 5):
            cv, nm = self.convs[k], self.norms[k]
            C = cv.Ca
            dbar = torch.empty_like(ctx.y[k])
            if cv.edge:
                ops.edge_down_img(abar, cv.w_down_e, dbar, None, C)
                nbytes = ops.edge_wgrad_workspace_bytes(B, cv.Hs, cv.Ws, C)
                ops.edge_wgrad_img(ctx.dy[k], abar, self._gb(cv.weight), self.ws.get(nbytes), C, self.nc, True)
            else:
                ops.conv_down(abar, cv.w_down, dbar, None, C, cv.Cb, algo=self.algo)
                nbytes = ops.wgrad_workspace_bytes(B, cv.Hs, cv.Ws, C, cv.Cb, self.dtype, self.algo)
                ops.conv_wgrad(ctx.dy[k], abar, self._gb(cv.weight), self.ws.get(nbytes), C, cv.Cb, True, algo=self.algo)
            asums = self.arena.take(3 * C) if self.arena is not None else None
            if asums is None:
                asums = torch.zeros(3 * C, dtype=torch.float32, device=self.dev)
            count = B * cv.Hs * cv.Ws * world
            ops.bn_adj_reduce(dbar, ctx.da[k], ctx.y[k], ctx.ss[k], ctx.mr[k], ctx.bsum[k], asums, C, count, LRELU)
            ops.bn_adj_param(asums, ctx.mr[k], self._gb(nm.bn.weight), C)          # this rank's share of dgamma
            self.comm.allreduce_sum_(asums)
            nxt, yb = torch.empty_like(ctx.y[k]), torch.empty_like(ctx.y[k])
            ops.bn_adj_apply(dbar, ctx.da[k], ctx.y[k], ctx.ss[k], ctx.mr[k], nm.gamma, ctx.bsum[k], asums, nxt, yb, C,
                             count, LRELU)
            abar, ybar[k] = nxt, yb
        hd = ctx.head
        ga4 = abar.view(B, self.F)
        gbar_h = self._f32(B, self.H)
        self._feat_fwd(ga4, gbar_h, B)
        self._feat_wgrad(hd["g_h"], ga4, B)
        gbar_hd = self._f32(B, self.H)
        ops.rowop(ops.ROW_MUL, gbar_h, hd["masks"], gbar_hd, B, self.H, s=KEEP_SCALE)
        w2 = self._p("linear2.weight")
        gbar_s = self._f32(B, 1)
        ops.dense(gbar_hd, self.H, 1, w2, self.H, 1, gbar_s, B, 1, self.H)
        ops.dense(hd["g_s"], 0, 1, gbar_hd, 1, self.H, self._gb(self.m.linear2.weight), 1, self.H, B, accumulate=True)
        sbar = self._f32(B)
        ops.logit_grad(ctx.prob, sbar, mode=3, up=gbar_s.view(-1))
        return sbar, ybar
