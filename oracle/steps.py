"""Oracle restatement of one G+D train step (TEST INFRASTRUCTURE, see oracle/__init__.py).

DCGAN: /root/reference/train/dcgan_trainer.py:155-189 (+ compute_gradient_penalty :110-127).
CGAN : /root/reference/train/cgan_trainer.py:173-213 (+ compute_gradient_penalty :114-131).

The reference draws its random tensors with torch.randn / torch.rand inside the step
(dcgan_trainer.py:160,168,171,111) and, for CGAN, inside nn.Dropout (CGAN.py:105).  CPU
mt19937 and CUDA Philox streams can never agree, so every random tensor is an *argument*
here (``rng``) -- the reference harness replays the same tensors into the unmodified
reference (oracle/ref_harness.py), which is how this file is pinned.

Every arithmetic expression keeps the reference's operand order (``0.9 * x + 0.1 * n``,
``alpha * real + ((1 - alpha) * fake)``, ``real + fake + lambda * gp``), so results are
bit-identical to the reference on the same torch build.
"""
import torch
from torch import nn

LABEL_REAL = 0.9      # dcgan_trainer.py:136
LABEL_FAKE = 0.1      # dcgan_trainer.py:137
LAMBDA_GP = 10.0      # dcgan_trainer.py:49
ADAM_BETAS = (0.5, 0.999)  # dcgan_trainer.py:61-62


def make_optimizers(g, d, lr):
    """dcgan_trainer.py:61-62 -- G first, then D."""
    opt_g = torch.optim.Adam(g.parameters(), lr=lr, betas=list(ADAM_BETAS))
    opt_d = torch.optim.Adam(d.parameters(), lr=lr, betas=list(ADAM_BETAS))
    return opt_g, opt_d


def make_rng(batch, nc=3, nz=100, hw=64, seed=0, n_steps=1, dropout_dim=None):
    """Deterministic per-step random tensors, in the order the reference draws them."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_steps):
        r = {
            "noise_real": torch.randn(batch, nc, hw, hw, generator=gen),   # :160
            "z": torch.randn(batch, nz, 1, 1, generator=gen),              # :168
            "noise_fake": torch.randn(batch, nc, hw, hw, generator=gen),   # :171
            "alpha": torch.rand(batch, 1, 1, 1, generator=gen),            # :111
        }
        if dropout_dim is not None:
            # four D passes per CGAN step, each with an independent keep-mask (p_drop = .25)
            r["drop"] = [(torch.rand(batch, dropout_dim, generator=gen) >= 0.25).float()
                         for _ in range(4)]
        out.append(r)
    return out


def make_real(batch, nc=3, hw=64, seed=12345, n_steps=1):
    """Synthetic 'real' images in [-1, 1): the range after Normalize(.5,.5)
    (dcgan_data_preprocessor.py:43)."""
    gen = torch.Generator().manual_seed(seed)
    return [torch.rand(batch, nc, hw, hw, generator=gen) * 2 - 1 for _ in range(n_steps)]


def gradient_penalty(d, real, fake, alpha, d_args=(), taps=None):
    """dcgan_trainer.py:110-127 / cgan_trainer.py:114-131."""
    x_hat = (alpha * real + ((1 - alpha) * fake)).requires_grad_(True)
    d_hat = d(x_hat, *d_args, taps=taps)
    grads = torch.autograd.grad(outputs=d_hat, inputs=x_hat,
                                grad_outputs=torch.ones_like(d_hat),
                                create_graph=True, retain_graph=True, only_inputs=True)[0]
    flat = grads.view(grads.size(0), -1)
    gp = ((flat.norm(2, dim=1) - 1) ** 2).mean()
    return gp, x_hat, d_hat, grads


def _param_grads(m):
    return {k: (p.grad.detach().clone() if p.grad is not None else None)
            for k, p in m.named_parameters()}


def dcgan_step(g, d, opt_g, opt_d, real, rng, capture=False):
    """One pass of dcgan_trainer.py:155-189.  Returns the scalars the reference logs
    (:191-196) and, with ``capture``, every intermediate a parity test wants."""
    bce = nn.BCELoss()                                    # :64
    out = {}
    cap = {} if capture else None

    # ---- D on real (A) :155-165
    d.zero_grad()
    b = real.size(0)
    label = torch.full((b,), LABEL_REAL, dtype=torch.float32)
    real_n = 0.9 * real + 0.1 * rng["noise_real"]
    taps_a = {} if capture else None
    p_real = d(real_n, taps=taps_a).view(-1)
    err_real = bce(p_real, label)
    err_real.backward()
    out["x_d"] = p_real.mean().item()

    # ---- G forward, D on fake.detach() (B) :168-176
    taps_g = {} if capture else None
    fake_raw = g(rng["z"], taps=taps_g)
    label.fill_(LABEL_FAKE)
    fake = 0.9 * fake_raw + 0.1 * rng["noise_fake"]
    taps_b = {} if capture else None
    p_fake = d(fake.detach(), taps=taps_b).view(-1)
    err_fake = bce(p_fake, label)
    err_fake.backward()
    out["z1_gd"] = p_fake.mean().item()

    # ---- gradient penalty (C) :178-180 -- logged, never back-propagated
    taps_c = {} if capture else None
    gp, x_hat, p_hat, gp_grads = gradient_penalty(d, real_n, fake, rng["alpha"], taps=taps_c)
    err_d = err_real + err_fake + LAMBDA_GP * gp
    if capture:
        cap["d_grads"] = _param_grads(d)
        cap["d_acts"] = {"A": {k: v.detach().clone() for k, v in taps_a.items() if not k.endswith('.pre')},
                         "B": {k: v.detach().clone() for k, v in taps_b.items() if not k.endswith('.pre')},
                         "C": {k: v.detach().clone() for k, v in taps_c.items() if not k.endswith('.pre')}}
        cap["d_act_grads"] = {"A": {k: v.grad.detach().clone() for k, v in taps_a.items() if not k.endswith('.pre')},
                              "B": {k: v.grad.detach().clone() for k, v in taps_b.items() if not k.endswith('.pre')}}
        cap["p_real"], cap["p_fake"], cap["p_hat"] = (p_real.detach().clone(), p_fake.detach().clone(),
                                                      p_hat.detach().view(-1).clone())
        cap["gp_grads"] = gp_grads.detach().clone()
        cap["fake_raw"] = fake_raw.detach().clone()
        cap["g_acts"] = {k: v.detach().clone() for k, v in taps_g.items() if not k.endswith('.pre')}
    opt_d.step()

    # ---- G step (D) :182-189
    g.zero_grad()
    label.fill_(LABEL_REAL)
    taps_d = {} if capture else None
    p_g = d(fake, taps=taps_d).view(-1)
    err_g = bce(p_g, label)
    err_g.backward()
    out["z2_gd"] = p_g.mean().item()
    if capture:
        cap["g_grads"] = _param_grads(g)
        cap["g_act_grads"] = {k: v.grad.detach().clone() for k, v in taps_g.items() if not k.endswith('.pre')}
        cap["d_acts"]["D"] = {k: v.detach().clone() for k, v in taps_d.items() if not k.endswith('.pre')}
        cap["d_act_grads"]["D"] = {k: v.grad.detach().clone() for k, v in taps_d.items() if not k.endswith('.pre')}
        cap["p_g"] = p_g.detach().clone()
        cap["pre"] = {tag: {k[:-4]: v for k, v in t.items() if k.endswith('.pre')}
                      for tag, t in (("A", taps_a), ("B", taps_b), ("C", taps_c), ("D", taps_d), ("G", taps_g))}
    opt_g.step()

    out.update(loss_d=err_d.item(), loss_g=err_g.item(), gp=gp.item(),
               err_real=err_real.item(), err_fake=err_fake.item())
    if capture:
        out["capture"] = cap
    return out


class _InjectedDropout(nn.Module):
    """nn.Dropout(p) with a caller-supplied keep-mask: y = x * mask / (1 - p), which is what
    torch's dropout computes for the mask it draws (CGAN.py:105,122)."""

    def __init__(self, p=0.25):
        super().__init__()
        self.p = p
        self.mask = None

    def forward(self, x):
        if not self.training or self.mask is None:
            return nn.functional.dropout(x, self.p, self.training)
        return x * self.mask / (1.0 - self.p)


def inject_dropout(d):
    """Swap CganDiscriminator.drop1 for the injectable variant (state_dict unaffected)."""
    if not isinstance(d.drop1, _InjectedDropout):
        d.drop1 = _InjectedDropout(d.drop1.p)
    return d


def cgan_step(g, d, opt_g, opt_d, real, labels, rng, capture=False):
    """One pass of cgan_trainer.py:173-213.  ``labels`` is the int64 one-hot [B,n_classes]
    the preprocessor yields (cgan_data_preprocessor.py:11-16)."""
    bce = nn.BCELoss()
    out = {}
    cap = {} if capture else None
    masks = rng.get("drop")

    def set_mask(i):
        if masks is not None:
            inject_dropout(d).drop1.mask = masks[i]

    d.zero_grad()
    b = real.size(0)
    label = torch.full((b,), LABEL_REAL, dtype=torch.float32)
    real_n = 0.9 * real + 0.1 * rng["noise_real"]                       # :182
    set_mask(0)
    taps = {t: ({} if capture else None) for t in "ABCDG"}
    p_real = d(real_n, labels.detach(), taps=taps["A"]).view(-1)        # :184
    err_real = bce(p_real, label)
    out["x_d"] = p_real.mean().item()

    fake_raw = g(rng["z"], labels.detach(), taps=taps["G"])             # :190
    label = torch.full((b,), LABEL_FAKE, dtype=torch.float32)           # :191
    fake = 0.9 * fake_raw + 0.1 * rng["noise_fake"]
    set_mask(1)
    p_fake = d(fake.detach(), labels.detach(), taps=taps["B"]).view(-1)  # :194
    err_fake = bce(p_fake, label)
    out["z1_gd"] = p_fake.mean().item()

    set_mask(2)
    gp, x_hat, p_hat, gp_grads = gradient_penalty(d, real_n.detach(), fake.detach(), rng["alpha"],
                                                  d_args=(labels.detach(),), taps=taps["C"])   # :200
    err_d = err_real.mean() + err_fake.mean() + LAMBDA_GP * gp          # :201
    err_d.backward()                                                    # :203 -- second order through D
    if capture:
        cap["d_grads"] = _param_grads(d)
        cap["p_real"], cap["p_fake"], cap["p_hat"] = (p_real.detach().clone(), p_fake.detach().clone(),
                                                      p_hat.detach().view(-1).clone())
        cap["gp_grads"] = gp_grads.detach().clone()
        cap["fake_raw"] = fake_raw.detach().clone()
    opt_d.step()

    g.zero_grad()
    label.fill_(LABEL_REAL)
    set_mask(3)
    p_g = d(fake, labels, taps=taps["D"]).view(-1)                      # :209
    err_g = bce(p_g, label)
    err_g.backward()
    out["z2_gd"] = p_g.mean().item()
    if capture:
        cap["g_grads"] = _param_grads(g)
        cap["p_g"] = p_g.detach().clone()
        cap["pre"] = {tag: {k[:-4]: v for k, v in t.items() if k.endswith('.pre')} for tag, t in taps.items()}
    opt_g.step()

    out.update(loss_d=err_d.item(), loss_g=err_g.item(), gp=gp.item(),
               err_real=err_real.item(), err_fake=err_fake.item())
    if capture:
        out["capture"] = cap
    return out


def one_hot(idx, n_classes):
    """cgan_data_preprocessor.py:11-16 (OneHotEncoder) for a batch of class indices."""
    return torch.nn.functional.one_hot(idx, n_classes).to(torch.int64)
