"""Drop-in for the reference's model/CGAN.py: `Generator`, `Discriminator`, `weights_init`.

Same constructors, attribute names and state_dict keys as /root/reference/model/CGAN.py:79-171
(label_embedding, label_embedding_relu1, conv1-4, norm1-4, relu1-4, flatten, linear1, drop1, linear2,
sigmoid / conv1-5, norm1-4, relu1-4, tanh); torch.nn layers are parameter containers, `forward` runs
the sm_100a kernels.  Additive kwargs: nc, nz, ngf / ndf, n_classes, embed, dtype.

`Discriminator.forward(x, labels)` is differentiable once (first order).  The CGAN *trainer* needs the
gradient penalty's second-order terms and gets them from engine_cgan.CganDiscriminatorEngine's explicit
sweep (train/cgan_step.py), not from torch.autograd.
"""
import torch
from torch import nn

from .. import ops
from ..engine import GeneratorEngine
from ..engine_cgan import CganDiscriminatorEngine, P_DROP
from .DCGAN import _require_cuda


def weights_init(m):
    # reference model/CGAN.py:165-171 -- matched on the class NAME, so Linear keeps torch's default init
    name = type(m).__name__
    if 'Conv' in name:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif 'BatchNorm' in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


class _CganDForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, labels, module, mask, *params):
        eng = module.engine()
        B = x.shape[0]
        x_img = ops.img_alloc(B, x.shape[1], x.shape[2], x.shape[3], eng.dtype, x.device, eng.img_layout)
        ops.prep_image(x.detach().contiguous().float(), out_nhwc=x_img, layout=eng.img_layout)
        c = eng.trunk_forward(x_img, groups=1, update_running=module.training)
        prob = eng.head_forward(c, labels, mask)
        ctx.c, ctx.module, ctx.x_shape = c, module, x.shape
        return prob.view(B, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dprob):
        module, c = ctx.module, ctx.c
        eng = module.engine()
        params = list(module.parameters())
        need_w = any(ctx.needs_input_grad[4:])
        need_x = ctx.needs_input_grad[0]
        B = c.B
        eng.sink = {}
        try:
            dls = torch.empty(B, dtype=torch.float32, device=dprob.device)
            ops.logit_grad(c.prob, dls, mode=2, up=dprob.detach().reshape(-1).contiguous().float())
            da4 = eng.head_backward(c, dls, wgrad=need_w)
            if need_w:
                eng.flush_linear1_grad(accumulate=False)
            dx_img = eng.trunk_backward(c, da4, wgrad=need_w, input_grad=need_x, accumulate=False)
            sink = eng.sink
        finally:
            eng.sink = None
        dx = None
        if need_x:
            dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dprob.device)
            ops.nhwc_to_nchw(dx_img, dx, layout=eng.img_layout)
        return (dx, None, None, None, *[sink.get(id(p)) if need_w else None for p in params])


class Discriminator(nn.Module):
    def __init__(self, nc=3, ndf=64, n_classes=100, embed=200, dtype=torch.bfloat16):
        super().__init__()
        self.label_embedding = nn.Linear(n_classes, embed)
        self.label_embedding_relu1 = nn.LeakyReLU(0.2, inplace=True)
        w = [nc, ndf, ndf * 2, ndf * 4, ndf * 8]
        for i in range(4):
            setattr(self, f"conv{i + 1}", nn.Conv2d(w[i], w[i + 1], kernel_size=4, stride=2, padding=1, bias=False))
            setattr(self, f"norm{i + 1}", nn.BatchNorm2d(w[i + 1]))
            setattr(self, f"relu{i + 1}", nn.LeakyReLU(0.2, inplace=True))
        self.flatten = nn.Flatten()
        self.linear1 = nn.Linear(w[4] * 16 + embed, 256)
        self.drop1 = nn.Dropout(P_DROP)
        self.linear2 = nn.Linear(256, 1)
        self.sigmoid = nn.Sigmoid()
        self.compute_dtype = dtype
        self._engine = None
        self._comm = None

    def set_compute(self, dtype=None, comm=None):
        if dtype is not None:
            self.compute_dtype = dtype
        if comm is not None:
            self._comm = comm
        self._engine = None
        return self

    def engine(self):
        dev = self.conv1.weight.device
        if self._engine is None or self._engine.dev != dev:
            if dev.type != "cuda":
                raise RuntimeError("Discriminator: parameters must live on a CUDA device (no CPU fallback)")
            self._engine = CganDiscriminatorEngine(self, self.compute_dtype, comm=self._comm)
        return self._engine

    def forward(self, x, labels, dropout_mask=None):
        """dropout_mask: optional [B,256] keep-mask (parity tests inject it); otherwise drawn on the device in
        train mode, all-ones in eval mode."""
        _require_cuda(x, "Discriminator.forward")
        B = x.shape[0]
        if dropout_mask is None:
            dropout_mask = torch.ones(B, 256, dtype=torch.float32, device=x.device)
            if self.training:
                if getattr(self, "_mask_ctr", None) is None:
                    self._mask_ctr = torch.zeros(1, dtype=torch.int64, device=x.device)
                ops.dropout_mask(dropout_mask, P_DROP, 12345, 77, self._mask_ctr)
        return _CganDForward.apply(x, labels, self, dropout_mask.contiguous().float(), *self.parameters())


class _CganGForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, module, *params):
        eng = module.engine()
        B = z.shape[0]
        z2d = module.concat_inputs(z, labels)
        c = eng.forward(z2d, update_running=module.training)
        out = torch.empty(B, eng.nc, 64, 64, dtype=torch.float32, device=z.device)
        ops.g_out_fwd(c.y[5], None, 1.0, 0.0, out, None, None, (B, eng.nc, 64, 64), layout=eng.img_layout)
        ctx.c, ctx.module, ctx.out = c, module, out
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        module, c = ctx.module, ctx.c
        eng = module.engine()
        params = list(module.parameters())
        B = dout.shape[0]
        lay = eng.img_layout
        d_img = ops.img_alloc(B, eng.nc, 64, 64, eng.dtype, dout.device, lay)
        ops.prep_image(dout.detach().contiguous().float(), out_nhwc=d_img, layout=lay)
        dy5 = torch.zeros_like(d_img) if lay == ops.IMG_P4 else torch.empty_like(d_img)
        ops.g_out_bwd(d_img, ctx.out, 1.0, dy5, layout=lay)
        eng.sink = {}
        try:
            eng.backward(c, dy5, accumulate=False)
            sink = eng.sink
        finally:
            eng.sink = None
        return (None, None, None, *[sink.get(id(p)) for p in params])


class Generator(nn.Module):
    def __init__(self, nc=3, nz=100, ngf=64, n_classes=100, dtype=torch.bfloat16):
        super().__init__()
        self.nz, self.n_classes = nz, n_classes
        w = [nz + n_classes, ngf * 8, ngf * 4, ngf * 2, ngf]
        for i in range(4):
            stride, pad = (1, 0) if i == 0 else (2, 1)
            setattr(self, f"conv{i + 1}", nn.ConvTranspose2d(w[i], w[i + 1], kernel_size=4, stride=stride,
                                                             padding=pad, bias=False))
            setattr(self, f"norm{i + 1}", nn.BatchNorm2d(w[i + 1]))
            setattr(self, f"relu{i + 1}", nn.ReLU(inplace=True))
        self.conv5 = nn.ConvTranspose2d(ngf, nc, kernel_size=4, stride=2, padding=1, bias=False)
        self.tanh = nn.Tanh()
        self.compute_dtype = dtype
        self._engine = None
        self._comm = None

    set_compute = Discriminator.set_compute

    def engine(self):
        dev = self.conv1.weight.device
        if self._engine is None or self._engine.dev != dev:
            if dev.type != "cuda":
                raise RuntimeError("Generator: parameters must live on a CUDA device (no CPU fallback)")
            self._engine = GeneratorEngine(self, self.compute_dtype, comm=self._comm)
        return self._engine

    def concat_inputs(self, z, labels):
        """cat([z, labels.reshape(-1, n_classes, 1, 1)], 1) of the reference (CGAN.py:154-155) as a [B, nz + n_classes]
        fp32 matrix, by one kernel (jck_concat_rows: concat + int64 -> float).  The trainer's step does not even materialise it:
        GeneratorEngine.forward(z, labels=...) writes the concatenation straight into conv1's bf16 operand."""
        B = z.shape[0]
        out = torch.empty(B, self.nz + self.n_classes, dtype=torch.float32, device=z.device)
        lab = labels.reshape(B, self.n_classes)
        if lab.dtype not in (torch.float32, torch.int64):
            lab = lab.float()
        ops.concat_rows(z.detach().reshape(B, self.nz).float().contiguous(), lab.contiguous(), out)    # one kernel
        return out

    def forward(self, x, labels):
        _require_cuda(x, "Generator.forward")
        return _CganGForward.apply(x, labels, self, *self.parameters())
