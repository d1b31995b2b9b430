"""Oracle restatement of the reference networks (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows /root/reference/model/DCGAN.py:6-76 and /root/reference/model/CGAN.py:79-171.
Attribute names (conv1..5, norm1..4, relu1..4, tanh / sigmoid, label_embedding,
linear1, drop1, linear2) and construction ORDER are the reference's, because
(a) ``state_dict`` keys must interchange and (b) the default initialisers consume
the global RNG stream in construction order, so seeding + ``weights_init`` only
reproduces the reference's weights if modules are created in the same sequence.

The literals the reference hard-codes are keyword arguments here:
    nc=3 (DCGAN.py:10,58)  nz=100 (DCGAN.py:42)  ngf=ndf=64
    n_classes=100, embed=200 (CGAN.py:83,104,132)
Defaults reproduce the reference exactly (asserted bit-for-bit in
tests/test_oracle_golden.py when /root/reference is present).
"""
import torch
from torch import nn


def _stack_names(n):
    return [(f"conv{i}", f"norm{i}", f"relu{i}") for i in range(1, n + 1)]


def _tap(taps, name, y):
    """Record a raw conv output (pre-BN; BN is out-of-place so the tensor stays intact even
    though the activation that follows BN is in-place) and keep its gradient."""
    if taps is not None:
        if y.requires_grad:
            y.retain_grad()
        taps[name] = y


def _bn_act(taps, name, norm, act, y):
    """BatchNorm then the (in-place) activation; with taps, also records the BatchNorm output -- the
    pre-activation whose sign decides the LeakyReLU / ReLU branch -- under ``name + ".pre"``."""
    pre = norm(y)
    if taps is not None:
        taps[name + ".pre"] = pre.detach().clone()
    return act(pre)


class DcganDiscriminator(nn.Module):
    """DCGAN.py:6-35 -- 4x (Conv k4 s2 p1 -> BN -> LeakyReLU .2), Conv k4 s1 p0, Sigmoid."""

    def __init__(self, nc=3, ndf=64):
        super().__init__()
        widths = [nc, ndf, ndf * 2, ndf * 4, ndf * 8]
        for i, (c, n, r) in enumerate(_stack_names(4)):
            setattr(self, c, nn.Conv2d(widths[i], widths[i + 1], 4, 2, 1, bias=False))
            setattr(self, n, nn.BatchNorm2d(widths[i + 1]))
            setattr(self, r, nn.LeakyReLU(0.2, inplace=True))
        self.conv5 = nn.Conv2d(widths[4], 1, 4, 1, 0, bias=False)
        self.sigmoid = nn.Sigmoid()

    def trunk(self, x, taps=None):
        h = x
        for c, n, r in _stack_names(4):
            y = getattr(self, c)(h)
            _tap(taps, c, y)
            h = _bn_act(taps, c, getattr(self, n), getattr(self, r), y)
        return h

    def forward(self, x, taps=None):
        return self.sigmoid(self.conv5(self.trunk(x, taps)))


class DcganGenerator(nn.Module):
    """DCGAN.py:38-67 -- ConvT(nz->8ngf, k4 s1 p0) then 3x ConvT k4 s2 p1 with BN+ReLU, ConvT -> Tanh."""

    def __init__(self, nc=3, nz=100, ngf=64, extra_in=0):
        super().__init__()
        widths = [nz + extra_in, ngf * 8, ngf * 4, ngf * 2, ngf]
        for i, (c, n, r) in enumerate(_stack_names(4)):
            stride, pad = (1, 0) if i == 0 else (2, 1)
            setattr(self, c, nn.ConvTranspose2d(widths[i], widths[i + 1], 4, stride, pad, bias=False))
            setattr(self, n, nn.BatchNorm2d(widths[i + 1]))
            setattr(self, r, nn.ReLU(inplace=True))
        self.conv5 = nn.ConvTranspose2d(ngf, nc, 4, 2, 1, bias=False)
        self.tanh = nn.Tanh()

    def forward(self, x, taps=None):
        h = x
        for c, n, r in _stack_names(4):
            y = getattr(self, c)(h)
            _tap(taps, c, y)
            h = _bn_act(taps, c, getattr(self, n), getattr(self, r), y)
        y = self.conv5(h)
        _tap(taps, "conv5", y)
        return self.tanh(y)


class CganDiscriminator(nn.Module):
    """CGAN.py:79-123 -- label Linear+LeakyReLU, DCGAN trunk (no conv5), flatten, cat,
    Linear(8ndf*16+embed -> 256), Dropout .25, Linear(256 -> 1), Sigmoid."""

    def __init__(self, nc=3, ndf=64, n_classes=100, embed=200):
        super().__init__()
        self.label_embedding = nn.Linear(n_classes, embed)
        self.label_embedding_relu1 = nn.LeakyReLU(0.2, inplace=True)
        widths = [nc, ndf, ndf * 2, ndf * 4, ndf * 8]
        for i, (c, n, r) in enumerate(_stack_names(4)):
            setattr(self, c, nn.Conv2d(widths[i], widths[i + 1], 4, 2, 1, bias=False))
            setattr(self, n, nn.BatchNorm2d(widths[i + 1]))
            setattr(self, r, nn.LeakyReLU(0.2, inplace=True))
        self.flatten = nn.Flatten()
        self.linear1 = nn.Linear(widths[4] * 16 + embed, 256)
        self.drop1 = nn.Dropout(0.25)
        self.linear2 = nn.Linear(256, 1)
        self.sigmoid = nn.Sigmoid()

    def forward(self, x, labels, taps=None):
        lab = self.label_embedding_relu1(self.label_embedding(labels.float()))
        h = x
        for c, n, r in _stack_names(4):
            y = getattr(self, c)(h)
            _tap(taps, c, y)
            h = _bn_act(taps, c, getattr(self, n), getattr(self, r), y)
        joined = torch.cat([self.flatten(h), lab], dim=1)
        return self.sigmoid(self.linear2(self.drop1(self.linear1(joined))))


class CganGenerator(DcganGenerator):
    """CGAN.py:127-162 -- one-hot labels reshaped [B,n_classes,1,1], concatenated to z."""

    def __init__(self, nc=3, nz=100, ngf=64, n_classes=100):
        super().__init__(nc=nc, nz=nz, ngf=ngf, extra_in=n_classes)
        self.n_classes = n_classes

    def forward(self, x, labels, taps=None):
        lab = labels.reshape(-1, self.n_classes, 1, 1)
        return super().forward(torch.cat([x, lab], 1), taps)


def weights_init(m):
    """DCGAN.py:70-76 / CGAN.py:165-171: Conv* ~ N(0,.02); BatchNorm gamma ~ N(1,.02), beta = 0.
    Class-name matching, as the reference does it: the oracle's classes are nn.Conv2d /
    nn.ConvTranspose2d / nn.BatchNorm2d instances, so the same names match; Linear keeps
    PyTorch's default init."""
    name = type(m).__name__
    if "Conv" in name:
        nn.init.normal_(m.weight.data, 0.0, 0.02)
    elif "BatchNorm" in name:
        nn.init.normal_(m.weight.data, 1.0, 0.02)
        nn.init.constant_(m.bias.data, 0)


def build(model="DCGAN", seed=12345, **kw):
    """Construct G then D and apply weights_init to G then D, the order of main.py:83-85 /
    dcgan_trainer.py:54-55, after seeding the global generator (main.py:34; RANDOMSEED=12345)."""
    torch.manual_seed(seed)
    if model == "DCGAN":
        g = DcganGenerator(**{k: v for k, v in kw.items() if k in ("nc", "nz", "ngf")})
        d = DcganDiscriminator(**{k: v for k, v in kw.items() if k in ("nc", "ndf")})
    else:
        g = CganGenerator(**{k: v for k, v in kw.items() if k in ("nc", "nz", "ngf", "n_classes")})
        d = CganDiscriminator(**{k: v for k, v in kw.items() if k in ("nc", "ndf", "n_classes", "embed")})
    g.apply(weights_init)
    d.apply(weights_init)
    return g, d
