"""Oracle multi-step runs and tensor digests (TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import torch

from . import models, steps


def digest(t, n=48):
    """Small, order-sensitive fingerprint of a tensor: float64 sum, L2 norm and a strided sample."""
    t = t.detach().to(torch.float64).flatten()
    stride = max(1, t.numel() // n)
    return {"numel": t.numel(), "sum": float(t.sum()), "l2": float(t.norm()),
            "sample": [float(v) for v in t[::stride][:n]]}


def digest_close(a, b, rtol, atol=0.0):
    """Compare two digests: relative on sum/l2 (scaled by l2), elementwise on the sample."""
    if a["numel"] != b["numel"]:
        return False, "numel"
    scale = max(abs(a["l2"]), abs(b["l2"]), 1e-30)
    if abs(a["l2"] - b["l2"]) > rtol * scale + atol:
        return False, f"l2 {a['l2']} vs {b['l2']}"
    sa, sb = torch.tensor(a["sample"]), torch.tensor(b["sample"])
    smax = max(float(sa.abs().max()), float(sb.abs().max()), 1e-30)
    err = float((sa - sb).abs().max())
    if err > rtol * smax + atol:
        return False, f"sample err {err} (max {smax})"
    return True, ""


def run_dcgan(real_batches, rng_steps, fixed_noise, lr, seed=12345, emulate_eval=True,
              capture_first=False, **kw):
    """Oracle counterpart of ref_harness.run_dcgan.  ``emulate_eval`` reproduces the one side
    effect the reference's eval branch (dcgan_trainer.py:198-221, fired at iters % 500 == 0 and on
    the last iteration) has on training state: G(fixed_noise) runs under no_grad but in train mode,
    which moves G's BatchNorm running statistics."""
    g, d = models.build("DCGAN", seed=seed, **kw)
    opt_g, opt_d = steps.make_optimizers(g, d, lr)
    losses_d, losses_g, first = [], [], None
    n = len(real_batches)
    for i, (real, rng) in enumerate(zip(real_batches, rng_steps)):
        out = steps.dcgan_step(g, d, opt_g, opt_d, real, rng, capture=(capture_first and i == 0))
        if i == 0:
            first = out
        losses_d.append(out["loss_d"])
        losses_g.append(out["loss_g"])
        if emulate_eval and (i % 500 == 0 or i == n - 1):
            with torch.no_grad():
                g(fixed_noise)
    return {"losses_d": losses_d, "losses_g": losses_g, "first": first,
            "g_state": {k: v.detach().clone() for k, v in g.state_dict().items()},
            "d_state": {k: v.detach().clone() for k, v in d.state_dict().items()},
            "opt_d": opt_d.state_dict(), "opt_g": opt_g.state_dict(), "g": g, "d": d}


def run_cgan(real_batches, label_batches, rng_steps, fixed_noise, fixed_labels, lr, seed=12345,
             emulate_eval=True, **kw):
    g, d = models.build("CGAN", seed=seed, **kw)
    steps.inject_dropout(d)
    opt_g, opt_d = steps.make_optimizers(g, d, lr)
    losses_d, losses_g = [], []
    n = len(real_batches)
    for i, (real, lab, rng) in enumerate(zip(real_batches, label_batches, rng_steps)):
        out = steps.cgan_step(g, d, opt_g, opt_d, real, lab, rng)
        losses_d.append(out["loss_d"])
        losses_g.append(out["loss_g"])
        if emulate_eval and (i % 500 == 0 or i == n - 1):
            with torch.no_grad():
                g(fixed_noise, fixed_labels)
    return {"losses_d": losses_d, "losses_g": losses_g,
            "g_state": {k: v.detach().clone() for k, v in g.state_dict().items()},
            "d_state": {k: v.detach().clone() for k, v in d.state_dict().items()}}
