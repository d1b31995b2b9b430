"""CPU proof of the second-order sweep used for the CGAN discriminator update (test infrastructure).

The CGAN trainer back-propagates the gradient penalty (train/cgan_trainer.py:200-204).  Instead of a
generic double-backward engine, the CUDA path runs: (1) the input-gradient sweep, (2) its ADJOINT from
the image side up to the head (conv forward / conv weight-gradient kernels + the BatchNorm-backward
adjoint + a head adjoint), (3) one ordinary backward sweep seeded with the logit second derivative and
with the adjoint's `ybar` terms injected at each raw conv output.  This script restates that algorithm
with plain torch CPU ops and checks it against torch.autograd's double backward on the oracle network.

    python -m tests.notes.cgan_second_order_check
"""
import torch
import torch.nn.functional as F

from oracle import models, steps

torch.manual_seed(0)
B, LAM, SLOPE, EPS = 6, 10.0, 0.2, 1e-5
g_, d = models.build("CGAN", seed=12345)
d = d.double()
steps.inject_dropout(d)
d.label_embedding.register_forward_pre_hook(lambda m, inp: (inp[0].double(),))   # the oracle casts labels .float()
x_hat = (torch.rand(B, 3, 64, 64, dtype=torch.double) * 2 - 1)
labels = steps.one_hot(torch.randint(0, 100, (B,)), 100)
mask = (torch.rand(B, 256) >= 0.25).double()
d.drop1.mask = mask

# ---------------- truth: autograd double backward -------------------------------------------------
d.zero_grad()
xr = x_hat.clone().requires_grad_(True)
p = d(xr, labels)
v = torch.autograd.grad(p, xr, torch.ones_like(p), create_graph=True)[0]
gp = LAM * ((v.view(B, -1).norm(2, dim=1) - 1) ** 2).mean()
gp.backward()
truth = {n: q.grad.clone() for n, q in d.named_parameters()}

# ---------------- the explicit algorithm -----------------------------------------------------------
W = [getattr(d, f"conv{k}").weight.detach() for k in range(1, 5)]
gam = [getattr(d, f"norm{k}").weight.detach() for k in range(1, 5)]
bet = [getattr(d, f"norm{k}").bias.detach() for k in range(1, 5)]
W1, b1 = d.linear1.weight.detach(), d.linear1.bias.detach()
w2, b2 = d.linear2.weight.detach()[0], d.linear2.bias.detach()
We, be = d.label_embedding.weight.detach(), d.label_embedding.bias.detach()
W1a, W1b = W1[:, :8192], W1[:, 8192:]
grads = {n: torch.zeros_like(q) for n, q in d.named_parameters()}


def bn_stats(y):
    mu = y.mean((0, 2, 3), keepdim=True)
    rs = (y.var((0, 2, 3), unbiased=False, keepdim=True) + EPS).rsqrt()
    return mu, rs


def bn_bwd(g, xh, c):          # c = gamma*rstd
    m = lambda t: t.mean((0, 2, 3), keepdim=True)
    return c * (g - m(g) - xh * m(g * xh))


with torch.no_grad():
    # forward
    a, ys, xhs, rss, ms = [x_hat], [], [], [], []
    for k in range(4):
        y = F.conv2d(a[-1], W[k], stride=2, padding=1)
        mu, rs = bn_stats(y)
        xh = (y - mu) * rs
        pre = xh * gam[k].view(1, -1, 1, 1) + bet[k].view(1, -1, 1, 1)
        ys.append(y); xhs.append(xh); rss.append(rs); ms.append(torch.where(pre > 0, 1.0, SLOPE))
        a.append(torch.where(pre > 0, pre, SLOPE * pre))
    e_pre = labels.double() @ We.t() + be
    e = torch.where(e_pre > 0, e_pre, SLOPE * e_pre)
    a4f = a[4].flatten(1)
    h = a4f @ W1a.t() + e @ W1b.t() + b1
    hd = h * mask / 0.75
    s = hd @ w2 + b2
    pr = torch.sigmoid(s)
    # (1) input-gradient sweep, upstream ones on the sigmoid output
    g_s = pr * (1 - pr)
    g_h = (g_s[:, None] * w2) * mask / 0.75
    g_a = [None] * 5
    g_a[4] = (g_h @ W1a).view_as(a[4])
    dk, N = [None] * 4, []
    for k in range(3, -1, -1):
        c = gam[k].view(1, -1, 1, 1) * rss[k]
        dk[k] = bn_bwd(g_a[k + 1] * ms[k], xhs[k], c)
        g_a[k] = F.conv_transpose2d(dk[k], W[k], stride=2, padding=1)
    vv = g_a[0]
    nrm = vv.view(B, -1).norm(2, dim=1)
    u = (LAM * 2 / B * (1 - 1 / nrm)).view(B, 1, 1, 1) * vv
    # (2) adjoint sweep, image side -> head
    abar, ybar = u, [None] * 4
    for k in range(4):
        dbar = F.conv2d(abar, W[k], stride=2, padding=1)
        grads[f"conv{k + 1}.weight"] += torch.nn.grad.conv2d_weight(abar, W[k].shape, dk[k], stride=2, padding=1)
        n = dbar.numel() / dbar.shape[1]
        sm = lambda t: t.sum((0, 2, 3), keepdim=True)
        c = gam[k].view(1, -1, 1, 1) * rss[k]
        g = g_a[k + 1] * ms[k]
        S1, S2 = sm(dbar), sm(dbar * xhs[k])
        sg, q = sm(g), sm(g * xhs[k]) / n
        r = g - sg / n - xhs[k] * q
        S3 = sm(dbar * r)
        gb = c * (dbar - S1 / n - xhs[k] * S2 / n)
        grads[f"norm{k + 1}.weight"] += (S3 * rss[k]).flatten()
        xb = -c * (q * dbar + S2 / n * g)
        T1, T2 = -c * (q * S1 + S2 / n * sg), -2 * c * q * S2
        ybar[k] = rss[k] * (xb - T1 / n - xhs[k] * T2 / n) - gam[k].view(1, -1, 1, 1) * S3 * rss[k] ** 2 / n * xhs[k]
        abar = gb * ms[k]
    gbar_h = abar.flatten(1) @ W1a.t()
    grads["linear1.weight"][:, :8192] += g_h.t() @ abar.flatten(1)
    gbar_hd = gbar_h * mask / 0.75
    gbar_s = gbar_hd @ w2
    grads["linear2.weight"] += (g_s[:, None] * gbar_hd).sum(0, keepdim=True)
    sbar = gbar_s * pr * (1 - pr) * (1 - 2 * pr)
    # (3) ordinary backward seeded with sbar, ybar injected at the raw conv outputs
    grads["linear2.weight"] += (sbar[:, None] * hd).sum(0, keepdim=True)
    grads["linear2.bias"] += sbar.sum(0, keepdim=True)
    gh = (sbar[:, None] * w2) * mask / 0.75
    grads["linear1.weight"] += gh.t() @ torch.cat([a4f, e], 1)
    grads["linear1.bias"] += gh.sum(0)
    ge = (gh @ W1b) * torch.where(e_pre > 0, 1.0, SLOPE)
    grads["label_embedding.weight"] += ge.t() @ labels.double()
    grads["label_embedding.bias"] += ge.sum(0)
    da = (gh @ W1a).view_as(a[4])
    for k in range(3, -1, -1):
        g = da * ms[k]
        grads[f"norm{k + 1}.weight"] += (g * xhs[k]).sum((0, 2, 3))
        grads[f"norm{k + 1}.bias"] += g.sum((0, 2, 3))
        dy = bn_bwd(g, xhs[k], gam[k].view(1, -1, 1, 1) * rss[k]) + ybar[k]
        grads[f"conv{k + 1}.weight"] += torch.nn.grad.conv2d_weight(a[k], W[k].shape, dy, stride=2, padding=1)
        if k > 0:
            da = F.conv_transpose2d(dy, W[k], stride=2, padding=1)

worst = 0.0
for n, t in truth.items():
    err = float((grads[n] - t).norm() / t.norm().clamp_min(1e-300))
    worst = max(worst, err)
    print(f"{err:.2e}  {n}")
assert worst < 1e-6, worst   # parameters are float32 values promoted to double
print("second-order sweep matches autograd double backward; gp =", float(gp))
