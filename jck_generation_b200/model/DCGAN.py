"""Drop-in for the reference's model/DCGAN.py: `Generator`, `Discriminator`, `weights_init`.

Same zero-argument constructors, attribute names (conv1..5, norm1..4, relu1..4, tanh / sigmoid) and
state_dict keys as /root/reference/model/DCGAN.py:6-76, so checkpoints interchange with the reference
in both directions and `.apply(weights_init)` / `.parameters()` / `.to(device)` keep working.  The
torch.nn layers are kept as *parameter containers* (constructed in the reference's order, so seeding
reproduces its initial weights); `forward` does not call them -- it runs the sm_100a kernels through
engine.py and is differentiable through a torch.autograd.Function whose backward is again our kernels.

Additive keyword arguments (defaults = the reference's literals): nc, nz, ngf / ndf, dtype.
`dtype=torch.bfloat16` selects the tcgen05 path, `torch.float32` the exact-parity CUDA-core path.
There is no CPU path: calling forward on CPU tensors raises.
"""
import torch
from torch import nn

from .. import ops
from ..engine import DiscriminatorEngine, GeneratorEngine


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who}: the B200 path has no CPU fallback; move the module and its inputs to "
                           "a CUDA device (the reference's CPU behaviour lives in oracle/ for tests only)")


class _DForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, *params):
        eng = module.engine()
        B = x.shape[0]
        x_nhwc = ops.img_alloc(B, x.shape[1], x.shape[2], x.shape[3], eng.dtype, x.device, eng.img_layout)
        ops.prep_image(x.detach().contiguous().float(), out_nhwc=x_nhwc, layout=eng.img_layout)
        c = eng.trunk_forward(x_nhwc, groups=1, update_running=module.training)
        prob = eng.head_forward(c)
        ctx.c, ctx.module = c, module
        ctx.x_shape = x.shape
        return prob.view(B, 1, 1, 1)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dprob):
        module, c = ctx.module, ctx.c
        eng = module.engine()
        params = list(module.parameters())
        need_w = any(ctx.needs_input_grad[2:])
        need_x = ctx.needs_input_grad[0]
        eng.sink = {}
        try:
            dp = dprob.detach().reshape(-1).contiguous().float()
            da4 = eng.head_backward(c, mode=2, dprob=dp, wgrad=need_w, accumulate=False)
            dx_nhwc = eng.trunk_backward(c, da4, wgrad=need_w, input_grad=need_x, accumulate=False)
            sink = eng.sink
        finally:
            eng.sink = None
        dx = None
        if need_x:
            dx = torch.empty(ctx.x_shape, dtype=torch.float32, device=dprob.device)
            ops.nhwc_to_nchw(dx_nhwc, dx, layout=eng.img_layout)
        grads = [sink.get(id(p)) if need_w else None for p in params]
        return (dx, None, *grads)


class Discriminator(nn.Module):
    def __init__(self, nc=3, ndf=64, dtype=torch.bfloat16):
        super().__init__()
        w = [nc, ndf, ndf * 2, ndf * 4, ndf * 8]
        for i in range(4):
            setattr(self, f"conv{i + 1}", nn.Conv2d(w[i], w[i + 1], kernel_size=4, stride=2, padding=1, bias=False))
            setattr(self, f"norm{i + 1}", nn.BatchNorm2d(w[i + 1]))
            setattr(self, f"relu{i + 1}", nn.LeakyReLU(0.2, inplace=True))
        self.conv5 = nn.Conv2d(w[4], 1, kernel_size=4, stride=1, padding=0, bias=False)
        self.sigmoid = nn.Sigmoid()
        self.compute_dtype = dtype
        self._engine = None
        self._comm = None

    def set_compute(self, dtype=None, comm=None):
        """Choose the arithmetic path (bf16 tcgen05 / fp32 CUDA-core) and the data-parallel communicator."""
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
            self._engine = DiscriminatorEngine(self, self.compute_dtype, comm=self._comm)
        return self._engine

    def forward(self, x):
        _require_cuda(x, "Discriminator.forward")
        return _DForward.apply(x, self, *self.parameters())


class _GForward(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, module, *params):
        eng = module.engine()
        B = z.shape[0]
        z2d = z.detach().reshape(B, -1).contiguous().float()
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
        d_nhwc = ops.img_alloc(B, eng.nc, 64, 64, eng.dtype, dout.device, lay)
        ops.prep_image(dout.detach().contiguous().float(), out_nhwc=d_nhwc, layout=lay)
        dy5 = torch.zeros_like(d_nhwc) if lay == ops.IMG_P4 else torch.empty_like(d_nhwc)
        ops.g_out_bwd(d_nhwc, ctx.out, 1.0, dy5, layout=lay)
        eng.sink = {}
        try:
            eng.backward(c, dy5, accumulate=False)
            sink = eng.sink
        finally:
            eng.sink = None
        return (None, None, *[sink.get(id(p)) for p in params])


class Generator(nn.Module):
    def __init__(self, nc=3, nz=100, ngf=64, dtype=torch.bfloat16):
        super().__init__()
        w = [nz, ngf * 8, ngf * 4, ngf * 2, ngf]
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

    def forward(self, x):
        _require_cuda(x, "Generator.forward")
        return _GForward.apply(x, self, *self.parameters())


def weights_init(m):
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)
