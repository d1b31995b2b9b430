"""Explicit forward / backward of the DCGAN generator and discriminator stacks on the sm_100a kernels.

This is the host side of the hot path: it sequences the C-ABI kernels (ops.py) for
    model/DCGAN.py:29-35   Discriminator.forward   (+ its autograd backward)
    model/DCGAN.py:61-67   Generator.forward       (+ its autograd backward)
without torch.autograd: every gradient kernel is ours, launched explicitly, so the whole train step is
a fixed launch sequence (CUDA-graph capturable).  The nn.Module mirrors (model/DCGAN.py) and the
trainers (train/*.py) are thin layers over the two engines here.

Data layout: activations NHWC in `dtype` (bf16 = tcgen05 path, fp32 = exact-parity CUDA-core path);
parameters stay the reference's fp32 nn.Parameters in the reference's layouts; packed copies of the
conv weights (GEMM operand layouts, activation dtype) are private caches refreshed after each update.

BatchNorm "groups": D is run on several independent batches at once (real / fake / interpolated), each
with its own batch statistics, exactly as the reference's separate D(...) calls -- the conv kernels
take the whole concatenation, the statistics are kept per group.
"""
import contextlib

import torch

from . import ops
from .parallel import LocalComm

LRELU = 0.2
# The input-gradient convolution can do the next BatchNorm-backward reduction in its epilogue (jck_conv_*_bnbwd).
# Measured at 512 images (tests/notes/conv_bench.py): the thread-per-row epilogue re-reads y under an L2 that the
# TMA stream already saturates, so it only beats the separate streaming pass where the K loop is long -- layers
# whose output has >= 256 channels (+3..6 us vs a 8..14 us reduce pass); at 64 / 128 channels it loses (+18..46 us).
FUSE_MIN_C = 256


def fuse_up_bnbwd(cv):
    """Does the input-gradient (up) convolution of layer `cv` run the BatchNorm-backward reduction of the layer below in
    its epilogue?  Only for >= 256 output channels (above).  The windowed 64-channel kernel (conv_up_win_kernel) has the
    fused epilogue too (jck_conv_up_bnbwd, parity-tested), but measured at 512 images it takes 61 us against 31 us for the
    plain kernel + 34 us for the streaming reduce pass it would replace: ~10 extra operations per element on the eight
    epilogue warps a 168-register budget allows make the epilogue, not the MMAs, the critical path (JCK_FUSE_WIN=1 to try)."""
    import os
    if cv.Cb == 64 and os.environ.get("JCK_FUSE_WIN", "0") == "1":
        return cv.Ca <= 128 and cv.Hs % 16 == 0 and cv.Ws % 16 == 0
    return cv.Cb >= FUSE_MIN_C


BN_EPS = 1e-5
BN_MOM = 0.1


class _Conv:
    """One 4x4 stride-2 layer: weight w4[Ca][Cb][4][4] + packed operand caches."""

    def __init__(self, weight, Hs, dtype, edge=False):
        self.weight = weight
        self.Ca, self.Cb = int(weight.shape[0]), int(weight.shape[1])
        self.Hs = self.Ws = Hs
        self.edge = edge          # image-side layer on the tcgen05 edge kernels (JCK_IMG_P4 image layout)
        dev = weight.device
        if edge:
            self.w_down_e = torch.empty(self.Ca * 64, dtype=dtype, device=dev)
            self.w_up9 = torch.empty(16 * 9 * self.Ca, dtype=dtype, device=dev)
        else:
            self.w_down = torch.empty(self.Ca * 16 * self.Cb, dtype=dtype, device=dev)
            self.w_up = torch.empty(16 * self.Cb * self.Ca, dtype=dtype, device=dev)
        self._seen = None

    def refresh(self, force=False):
        key = (self.weight._version, self.weight.data_ptr())
        if force or key != self._seen:
            if self.edge:
                ops.pack_weights_edge(self.weight.detach(), self.w_down_e, self.w_up9)
            else:
                ops.pack_weights(self.weight.detach(), self.w_down, self.w_up)
            self._seen = key


def use_edge_kernels(dtype, algo, Ca, nc):
    """The image-side layers (nc <= 4 channels) run on tcgen05 when the arithmetic is bf16."""
    return dtype == torch.bfloat16 and algo != ops.ALGO_SIMT and Ca == 64 and nc <= 4


def _edge_up(cv, x_small, img_p4):
    """image-edge up conv: scatter form where the rows are 32 pixels (the 64x64 image), else the 9-shift gather form"""
    if cv.Ws == 32:
        ops.edge_up_scatter(x_small, cv.w_down_e, img_p4, cv.Ca)
    else:
        ops.edge_up(x_small, cv.w_up9, img_p4, cv.Ca)


class _Norm:
    def __init__(self, bn):
        self.bn = bn
        self.C = bn.num_features

    @property
    def gamma(self):
        return self.bn.weight.detach()

    @property
    def beta(self):
        return self.bn.bias.detach()


class Ctx:
    """Saved tensors of one forward pass (raw conv outputs, activations, BN coefficients)."""

    def __init__(self):
        self.x = None          # network input (NHWC)
        self.y = {}            # raw conv outputs per layer
        self.a = {}            # activations per layer
        self.ss = {}           # BN scale/shift [groups][2C]
        self.mr = {}           # BN mean/rstd   [groups][2C]
        self.prob = None
        self.groups = 1
        self.B = 0
        self.dy = {}           # gradients w.r.t. the raw conv outputs, filled by the backward sweep
        self.da = {}           # gradients w.r.t. the activations (inputs of the BatchNorm backward)
        self.bsum = {}         # BatchNorm backward sums [groups][2C]
        self.head = {}         # CGAN head intermediates

    def slice(self, g0, g1):
        """View of groups [g0, g1) of a grouped pass."""
        per = self.B // self.groups
        s = Ctx()
        s.groups, s.B = g1 - g0, per * (g1 - g0)
        lo, hi = per * g0, per * g1
        s.x = self.x[lo:hi]
        s.y = {k: v[lo:hi] for k, v in self.y.items()}
        s.a = {k: v[lo:hi] for k, v in self.a.items()}
        s.ss = {k: v[g0:g1] for k, v in self.ss.items()}
        s.mr = {k: v[g0:g1] for k, v in self.mr.items()}
        s.prob = self.prob[lo:hi] if self.prob is not None else None
        return s


def _zero_blocks(groups, channels, device, arena=None):
    """All the [groups][2C] fp32 accumulators of a pass (stats / backward sums), zeroed: slices of the step's arena
    (ops.ZeroArena: one memset per STEP) when the step provides one, else one torch.zeros for the pass (module API)."""
    total = sum(groups * 2 * c for c in channels)
    flat = arena.take(total) if arena is not None else None
    if flat is None:
        flat = torch.zeros(total, dtype=torch.float32, device=device)
    out, off = [], 0
    for c in channels:
        out.append(flat[off:off + groups * 2 * c].view(groups, 2 * c))
        off += groups * 2 * c
    return out


def _bn_finalize(comm, stats, nm, update_running, ss, mr, C, groups, count):
    """Batch statistics over the GLOBAL batch -> scale/shift, mean/rstd, running-buffer update.  With a peer
    communicator the exchange happens inside the finalize kernel (one launch, no NCCL call)."""
    bn = nm.bn
    run = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if update_running else (None, None, None)
    if comm.peer is not None and stats.numel() <= ops.COMM_MAX_N:
        ops.bn_finalize_sync(comm.peer, stats, nm.gamma, nm.beta, *run, ss, mr, C, groups, count, BN_EPS, BN_MOM)
    else:
        comm.allreduce_sum_(stats)
        ops.bn_finalize(stats, nm.gamma, nm.beta, *run, ss, mr, C, groups, count, BN_EPS, BN_MOM)


def _bn_bwd_sums(comm, sums, dgamma, dbeta, C, groups, accumulate):
    """BatchNorm-backward sums: local dgamma / dbeta (when asked), then the sums over the global batch."""
    if comm.peer is not None and sums.numel() <= ops.COMM_MAX_N:
        ops.bn_bwd_sums_sync(comm.peer, sums, dgamma, dbeta, C, groups, accumulate)
    else:
        if dgamma is not None:
            ops.bn_param_grad(sums, dgamma, dbeta, C, groups, accumulate)
        comm.allreduce_sum_(sums)


class _Workspace:
    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() * 4 < nbytes:
            self.buf = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
        return self.buf


class _GradTarget:
    """Where parameter gradients are written: the parameters' own .grad buffers (trainer hot path) or,
    when `sink` is a dict, fresh tensors collected for torch.autograd (module API).

    `wgrad_stream`: when the step sets it, the weight-gradient kernels (tensor bound, not on the critical path
    of the sweep) are launched on that side stream, forked after the tensor they read is complete, so they overlap
    the HBM-bound BatchNorm-backward streams of the next layer.  Whoever sets it must call join_wgrad() before the
    gradients are used; every operand is kept alive by the Ctx until then."""
    sink = None
    wgrad_stream = None
    arena = None              # ops.ZeroArena of the running step (set by the step; None: accumulators come from torch.zeros)
    grad_sync = None          # parallel.GradBuckets: told as each parameter's gradient becomes final (trainer hot path)

    def _grad_ready(self, *params):
        if self.grad_sync is not None and self.sink is None:
            self.grad_sync.ready(*params)

    def _wgrad_scope(self):
        if self.wgrad_stream is None:
            return contextlib.nullcontext()
        self.wgrad_stream.wait_stream(torch.cuda.current_stream())
        self._wgrad_pending = True
        return torch.cuda.stream(self.wgrad_stream)

    def join_wgrad(self):
        if self.wgrad_stream is not None and getattr(self, "_wgrad_pending", False):
            torch.cuda.current_stream().wait_stream(self.wgrad_stream)
            self._wgrad_pending = False

    def _gb(self, p):
        if self.sink is not None:
            buf = self.sink.get(id(p))
            if buf is None:
                buf = self.sink[id(p)] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            return buf
        if p.grad is None:
            p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return p.grad


class DiscriminatorEngine(_GradTarget):
    """conv1..4 (down) + BN + LeakyReLU(0.2), conv5 + sigmoid head.  model/DCGAN.py:6-35."""

    def __init__(self, module, dtype=torch.bfloat16, comm=None, algo=ops.ALGO_AUTO):
        self.m = module
        self.dtype = dtype
        self.algo = algo
        self.comm = comm or LocalComm()
        dev = module.conv1.weight.device
        self.dev = dev
        self.nc = int(module.conv1.weight.shape[1])
        edge = use_edge_kernels(dtype, algo, int(module.conv1.weight.shape[0]), self.nc)
        self.img_layout = ops.IMG_P4 if edge else ops.IMG_NHWC
        self.convs = {k: _Conv(getattr(module, f"conv{k}").weight, 64 >> k, dtype, edge=(edge and k == 1))
                      for k in range(1, 5)}
        self.norms = {k: _Norm(getattr(module, f"norm{k}")) for k in range(1, 5)}
        self.fused_bn_bwd = dtype == torch.bfloat16 and algo != ops.ALGO_SIMT
        self.has_head = hasattr(module, "conv5")
        if self.has_head:
            self.K5 = 16 * self.convs[4].Ca
            self.w5 = torch.empty(self.K5, dtype=dtype, device=dev)
            self.dw5 = torch.zeros(self.K5, dtype=torch.float32, device=dev)
            self._w5_seen = None
        self.ws = _Workspace(dev)

    def refresh(self, force=False):
        for c in self.convs.values():
            c.refresh(force)
        if self.has_head:
            w = self.m.conv5.weight
            key = (w._version, w.data_ptr())
            if force or key != self._w5_seen:
                ops.pack_head(w.detach(), self.w5)
                self._w5_seen = key

    # ---- forward -----------------------------------------------------------------------------------
    def trunk_forward(self, x_nhwc, groups=1, update_running=True):
        self.refresh()
        B = x_nhwc.shape[0]
        assert B % groups == 0
        ctx = Ctx()
        ctx.x, ctx.groups, ctx.B = x_nhwc, groups, B
        cur = x_nhwc
        world = self.comm.world_size
        zeros = _zero_blocks(groups, [self.convs[k].Ca for k in range(1, 5)], self.dev, self.arena)
        for k in range(1, 5):
            cv, nm = self.convs[k], self.norms[k]
            y = torch.empty(B, cv.Hs, cv.Ws, cv.Ca, dtype=self.dtype, device=self.dev)
            stats = zeros[k - 1]
            if cv.edge:
                ops.edge_down_img(cur, cv.w_down_e, y, stats, cv.Ca, ipg=B // groups)   # patch matrix only if wgrad needs it
            else:
                ops.conv_down(cur, cv.w_down, y, stats, cv.Ca, cv.Cb, ipg=B // groups, algo=self.algo)
            ss = torch.empty(groups, 2 * cv.Ca, dtype=torch.float32, device=self.dev)
            mr = torch.empty(groups, 2 * cv.Ca, dtype=torch.float32, device=self.dev)
            count = (B // groups) * cv.Hs * cv.Ws * world
            _bn_finalize(self.comm, stats, nm, update_running, ss, mr, cv.Ca, groups, count)   # SyncBN: global batch statistics
            a = torch.empty_like(y)
            ops.bn_act_fwd(y, ss, a, cv.Ca, groups, LRELU)
            ctx.y[k], ctx.a[k], ctx.ss[k], ctx.mr[k] = y, a, ss, mr
            cur = a
        return ctx

    def head_forward(self, ctx, targets=None, scalars=None):
        """prob = sigmoid(conv5(a4)); with `targets` (one per group) also accumulates, per group g,
        scalars[g][0] += BCE mean and scalars[g][1] += mean(prob)."""
        B, per = ctx.B, ctx.B // ctx.groups
        prob = torch.empty(B, dtype=torch.float32, device=self.dev)
        a4 = ctx.a[4].view(B, self.K5)
        for g in range(ctx.groups):
            sc = scalars[g] if (scalars is not None and targets is not None and targets[g] is not None) else None
            t = targets[g] if (targets is not None and targets[g] is not None) else 0.0
            ops.head_fwd(a4[g * per:(g + 1) * per], self.w5, prob[g * per:(g + 1) * per], t, sc)
        ctx.prob = prob
        return prob

    # ---- backward ----------------------------------------------------------------------------------
    def head_backward(self, ctx, mode, targets=None, dprob=None, wgrad=True, accumulate=False):
        """d(loss)/d(a4) (+ conv5 weight gradient).  mode 0 BCE-mean, 1 ones (GP sweep), 2 upstream dprob.
        The BCE mean runs over the GLOBAL batch (rows per group x world size): under data parallelism the ranks'
        gradients then SUM to the gradient of the reference's global-batch mean -- no division after the exchange."""
        B, per = ctx.B, ctx.B // ctx.groups
        da4 = torch.empty(B, self.K5, dtype=self.dtype, device=self.dev)
        a4 = ctx.a[4].view(B, self.K5)
        if wgrad:
            ops.zero(self.dw5)
        for g in range(ctx.groups):
            sl = slice(g * per, (g + 1) * per)
            ops.head_bwd(ctx.prob[sl], targets[g] if targets is not None else 0.0, self.w5, a4[sl], da4[sl],
                         self.dw5 if wgrad else None, mode, True,
                         dprob=dprob[sl] if dprob is not None else None, mean_count=per * self.comm.world_size)
        if wgrad:
            ops.unpack_head_grad(self.dw5, self._gb(self.m.conv5.weight), accumulate)
            if not accumulate:
                self._grad_ready(self.m.conv5.weight)
        return da4.view(B, 4, 4, self.convs[4].Ca)

    def trunk_backward(self, ctx, da4, wgrad=True, input_grad=False, accumulate=False, inject=None, inject_rows=None,
                       fuse=True, dx_out=None, comm=None):
        """Backward through conv4..conv1 given d/d(a4).  Returns d/d(input) (NHWC) when asked.
        `inject[k]` (rows `inject_rows` of the batch) is added to the gradient of the raw conv-k output
        before it is used: the second-order terms of the CGAN gradient penalty enter here.  The sweep
        records ctx.dy[k] and ctx.bsum[k] (BatchNorm backward sums).
        `fuse` (bf16 / tcgen05): the input-gradient convolution of layer k also performs the BatchNorm-backward
        reduction of layer k-1 in its epilogue and hands down g = da * act'(pre) instead of da; with
        fuse=False every layer runs the separate reduce pass and ctx.da[k] keeps d/d(activation) (the CGAN
        penalty sweep needs it).
        `dx_out`: a caller-owned JCK_IMG_P4 buffer (zero border / pad channel, e.g. one kept across steps) that
        receives the image-side input gradient instead of a freshly zeroed one.
        `comm`: communicator for this sweep's SyncBN exchanges when it is issued from a second stream
        (parallel.AuxComm); default the engine's."""
        B, groups = ctx.B, ctx.groups
        comm = comm or self.comm
        world = comm.world_size
        fuse = fuse and self.fused_bn_bwd
        da, reduced = da4, False
        zeros = _zero_blocks(groups, [self.convs[k].Ca for k in range(1, 5)], self.dev, self.arena)
        for k in range(4, 0, -1):
            cv, nm = self.convs[k], self.norms[k]
            C = cv.Ca
            sums = zeros[k - 1]
            if not reduced:
                ops.bn_act_bwd_reduce(da, ctx.y[k], ctx.ss[k], ctx.mr[k], sums, C, groups, LRELU)
            # parameter gradients are this rank's contribution (ranks are averaged later); then the global sums
            _bn_bwd_sums(comm, sums, self._gb(nm.bn.weight) if wgrad else None, self._gb(nm.bn.bias) if wgrad else None,
                         C, groups, accumulate)
            if wgrad and not accumulate:
                self._grad_ready(nm.bn.weight, nm.bn.bias)
            dy = torch.empty_like(ctx.y[k])
            count = (B // groups) * cv.Hs * cv.Ws * world
            # traversal: ascending behind the (descending) reduce pass, descending behind a fused convolution (ops.ORDER_*)
            ops.bn_act_bwd_apply(da, ctx.y[k], ctx.ss[k], ctx.mr[k], nm.gamma, sums, dy, C, groups, count,
                                 1.0 if reduced else LRELU,     # a fused producer already applied act'
                                 order=ops.ORDER_DESC if reduced else ops.ORDER_ASC)
            ctx.dy[k], ctx.bsum[k] = dy, sums
            ctx.da[k] = None if reduced else da
            if inject is not None:
                lo, hi = inject_rows
                ops.axpy(inject[k], dy[lo:hi], 1.0)
            inp = ctx.a[k - 1] if k > 1 else ctx.x
            if wgrad:
                with self._wgrad_scope():
                    if cv.edge:
                        nbytes = ops.edge_wgrad_workspace_bytes(B, cv.Hs, cv.Ws, cv.Ca)
                        ops.edge_wgrad_img(dy, ctx.x, self._gb(cv.weight), self.ws.get(nbytes), cv.Ca, self.nc, accumulate)
                    else:
                        nbytes = ops.wgrad_workspace_bytes(B, cv.Hs, cv.Ws, cv.Ca, cv.Cb, self.dtype, self.algo)
                        ops.conv_wgrad(dy, inp, self._gb(cv.weight), self.ws.get(nbytes), cv.Ca, cv.Cb, accumulate,
                                       algo=self.algo)
                    if not accumulate:
                        self._grad_ready(cv.weight)
            reduced = False
            if k > 1 or input_grad:
                if cv.edge:
                    # border / pad channel of the P4 image stay zero (the kernel writes interior pixels only)
                    da = dx_out if dx_out is not None else torch.zeros_like(inp)
                    _edge_up(cv, dy, da)
                elif fuse and k > 1 and fuse_up_bnbwd(cv):
                    da = torch.empty_like(inp)
                    ops.conv_up_bnbwd(dy, cv.w_up, ctx.y[k - 1], ctx.ss[k - 1], ctx.mr[k - 1], LRELU, da, zeros[k - 2],
                                      cv.Ca, cv.Cb, ipg=B // groups)
                    reduced = True
                else:
                    da = torch.empty_like(inp)
                    ops.conv_up(dy, cv.w_up, da, None, cv.Ca, cv.Cb, algo=self.algo)
            else:
                da = None
        return da


class GeneratorEngine(_GradTarget):
    """conv1 (dense 1x1 -> 4x4) + 3x (BN, ReLU, up) + conv5 up + tanh.  model/DCGAN.py:38-67."""

    def __init__(self, module, dtype=torch.bfloat16, comm=None, algo=ops.ALGO_AUTO):
        self.m = module
        self.dtype = dtype
        self.algo = algo
        self.comm = comm or LocalComm()
        dev = module.conv1.weight.device
        self.dev = dev
        w1 = module.conv1.weight
        self.K1, self.C1 = int(w1.shape[0]), int(w1.shape[1])
        # conv1 (1x1 -> 4x4) is a matrix product: on tcgen05 (jck_gemm_tc) in bf16 mode -- weight kept as the MN-major
        # operand [K1][16*C1], z cast to bf16 rows of pitch K1p -- and on the exact CUDA-core kernel in fp32 mode
        self.tc_fc = dtype == torch.bfloat16 and algo != ops.ALGO_SIMT
        self.K1p = (self.K1 + 7) // 8 * 8
        if self.tc_fc:
            self.w_fc = torch.empty(self.K1, 16 * self.C1, dtype=dtype, device=dev)
            self.dw_fc = torch.empty(self.K1, 16 * self.C1, dtype=torch.float32, device=dev)
        else:
            self.w_fc = torch.empty(16 * self.C1, self.K1, dtype=dtype, device=dev)
            self.dw_fc = torch.empty(16 * self.C1, self.K1, dtype=torch.float32, device=dev)
        self._w1_seen = None
        w5 = module.conv5.weight
        self.nc = int(w5.shape[1])
        edge = use_edge_kernels(dtype, algo, int(w5.shape[0]), self.nc)
        self.img_layout = ops.IMG_P4 if edge else ops.IMG_NHWC
        self.convs = {k: _Conv(getattr(module, f"conv{k}").weight, 2 << (k - 1), dtype, edge=(edge and k == 5))
                      for k in range(2, 6)}
        self.norms = {k: _Norm(getattr(module, f"norm{k}")) for k in range(1, 5)}
        self.ws = _Workspace(dev)
        self.gws = _Workspace(dev)

    def refresh(self, force=False):
        w = self.m.conv1.weight
        key = (w._version, w.data_ptr())
        if force or key != self._w1_seen:
            if self.tc_fc:
                ops.pack_fc_t(w.detach(), self.w_fc)
            else:
                ops.pack_fc(w.detach(), self.w_fc)
            self._w1_seen = key
        for c in self.convs.values():
            c.refresh(force)

    def _bn_relu(self, ctx, k, y, stats, groups, update_running):
        nm = self.norms[k]
        C = nm.C
        ss = torch.empty(groups, 2 * C, dtype=torch.float32, device=self.dev)
        mr = torch.empty(groups, 2 * C, dtype=torch.float32, device=self.dev)
        count = (y.numel() // C // groups) * self.comm.world_size
        _bn_finalize(self.comm, stats, nm, update_running, ss, mr, C, groups, count)
        a = torch.empty_like(y)
        ops.bn_act_fwd(y, ss, a, C, groups, 0.0)
        ctx.y[k], ctx.a[k], ctx.ss[k], ctx.mr[k] = y, a, ss, mr
        return a

    def forward(self, z2d, update_running=True, y5_out=None, labels=None):
        """z2d: [B, K1] fp32 (z, or cat(z, one-hot) for CGAN).  Returns ctx; ctx.y[5] is the raw conv5
        output [B,64,64,nc] (tanh is applied by ops.g_out_fwd at the image edge).  `y5_out`: caller-owned
        JCK_IMG_P4 buffer for it (zero border / pad channel), else a freshly zeroed one.
        `labels` ([B, n_classes] fp32 or int64, CGAN): z2d then holds the nz noise columns only and the reference's
        cat([z, labels], 1) (model/CGAN.py:154-155) is written straight into conv1's operand by ONE kernel
        (jck_concat_rows: concat + int64 -> float + fp32 -> bf16), not materialised by torch.cat."""
        self.refresh()
        B = z2d.shape[0]
        ctx = Ctx()
        zb_ready = False
        if labels is not None:
            assert z2d.shape[1] + labels.shape[1] == self.K1
            if self.tc_fc:
                ctx.zb = torch.empty(B, self.K1p, dtype=torch.bfloat16, device=self.dev)
                ops.concat_rows(z2d.contiguous(), labels.contiguous(), ctx.zb)
                zb_ready = True
            else:
                full = torch.empty(B, self.K1, dtype=torch.float32, device=self.dev)
                ops.concat_rows(z2d.contiguous(), labels.contiguous(), full)
                z2d = full
        ctx.x, ctx.B, ctx.groups = z2d, B, 1
        y1 = torch.empty(B, 4, 4, self.C1, dtype=self.dtype, device=self.dev)
        zeros = _zero_blocks(1, [self.norms[k].C for k in range(1, 5)], self.dev, self.arena)
        stats = zeros[0]
        N1 = 16 * self.C1
        if self.tc_fc:
            if not zb_ready:
                ctx.zb = torch.empty(B, self.K1p, dtype=torch.bfloat16, device=self.dev)
                ops.cast_rows_bf16(z2d.contiguous(), ctx.zb)
            ops.gemm_tc(ctx.zb, 0, self.K1p, self.w_fc, 1, N1, y1.view(B, N1), B, N1, self.K1, stats=stats, stats_channels=self.C1)
        else:
            ops.fc_fwd(z2d, self.w_fc, y1, stats, self.C1)
        cur = self._bn_relu(ctx, 1, y1, stats, 1, update_running)
        for k in range(2, 6):
            cv = self.convs[k]
            if cv.edge:
                y = y5_out if y5_out is not None else ops.img_alloc(B, self.nc, 2 * cv.Hs, 2 * cv.Ws, self.dtype, self.dev,
                                                                    ops.IMG_P4)
                _edge_up(cv, cur, y)
                ctx.y[5] = y
                continue
            y = torch.empty(B, 2 * cv.Hs, 2 * cv.Ws, cv.Cb, dtype=self.dtype, device=self.dev)
            if k < 5:
                stats = zeros[k - 1]
                ops.conv_up(cur, cv.w_up, y, stats, cv.Ca, cv.Cb, algo=self.algo)
                cur = self._bn_relu(ctx, k, y, stats, 1, update_running)
            else:
                ops.conv_up(cur, cv.w_up, y, None, cv.Ca, cv.Cb, algo=self.algo)
                ctx.y[5] = y
        return ctx

    def backward(self, ctx, dy5, accumulate=False):
        """dy5: gradient w.r.t. the raw conv5 output (NHWC).  Fills .grad of every generator parameter."""
        B = ctx.B
        world = self.comm.world_size
        d_large = dy5
        zeros = _zero_blocks(1, [self.norms[k].C for k in range(1, 5)], self.dev, self.arena)
        fuse = self.dtype == torch.bfloat16 and self.algo != ops.ALGO_SIMT
        for k in range(5, 1, -1):
            cv = self.convs[k]
            nm = self.norms[k - 1]
            C = nm.C
            sums = zeros[k - 2]
            yk, ssk, mrk = ctx.y[k - 1], ctx.ss[k - 1], ctx.mr[k - 1]
            da = torch.empty_like(ctx.a[k - 1])
            if cv.edge:
                nbytes = ops.edge_wgrad_workspace_bytes(B, cv.Hs, cv.Ws, cv.Ca)
                with self._wgrad_scope():
                    ops.edge_wgrad_img(ctx.a[k - 1], d_large, self._gb(cv.weight), self.ws.get(nbytes), cv.Ca, self.nc,
                                       accumulate)
                    if not accumulate:
                        self._grad_ready(cv.weight)
                ops.edge_down_img(d_large, cv.w_down_e, da, None, cv.Ca)
                reduced = False
            else:
                nbytes = ops.wgrad_workspace_bytes(B, cv.Hs, cv.Ws, cv.Ca, cv.Cb, self.dtype, self.algo)
                with self._wgrad_scope():
                    ops.conv_wgrad(ctx.a[k - 1], d_large, self._gb(cv.weight), self.ws.get(nbytes), cv.Ca, cv.Cb,
                                   accumulate, algo=self.algo)
                    if not accumulate:
                        self._grad_ready(cv.weight)
                reduced = fuse and cv.Ca >= FUSE_MIN_C
                if reduced:
                    ops.conv_down_bnbwd(d_large, cv.w_down, yk, ssk, mrk, 0.0, da, sums, cv.Ca, cv.Cb)
                else:
                    ops.conv_down(d_large, cv.w_down, da, None, cv.Ca, cv.Cb, algo=self.algo)
            if not reduced:
                ops.bn_act_bwd_reduce(da, yk, ssk, mrk, sums, C, 1, 0.0)
            _bn_bwd_sums(self.comm, sums, self._gb(nm.bn.weight), self._gb(nm.bn.bias), C, 1, accumulate)
            if not accumulate:
                self._grad_ready(nm.bn.weight, nm.bn.bias)
            dy = torch.empty_like(yk)
            count = (yk.numel() // C) * world
            # a fused producer already applied relu': slope 1 leaves g untouched
            ops.bn_act_bwd_apply(da, yk, ssk, mrk, nm.gamma, sums, dy, C, 1, count, 1.0 if reduced else 0.0,
                                 order=ops.ORDER_DESC if reduced else ops.ORDER_ASC)
            ctx.dy[k - 1] = dy
            d_large = dy
        N1 = 16 * self.C1
        if self.tc_fc:
            # dw_t[k][n] = sum_b z[b][k] * dy[b][n]: both operands MN-major, split over the batch
            nbytes = ops.gemm_tc_workspace_bytes(self.K1, N1, B)
            ops.gemm_tc(ctx.zb, 1, self.K1p, d_large.view(B, N1), 1, N1, self.dw_fc, self.K1, N1, B,
                        workspace=self.gws.get(nbytes) if nbytes else None)   # self.ws belongs to the side-stream wgrads
            ops.unpack_fc_grad_t(self.dw_fc, self._gb(self.m.conv1.weight), accumulate)
        else:
            ops.fc_wgrad(d_large.view(B, N1), ctx.x, self.dw_fc, accumulate=False)
            ops.unpack_fc_grad(self.dw_fc, self._gb(self.m.conv1.weight), accumulate)
        if not accumulate:
            self._grad_ready(self.m.conv1.weight)
