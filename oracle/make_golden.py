"""Freeze digests of the REFERENCE's own outputs into tests/golden/ (TEST INFRASTRUCTURE).

Run in the build container (needs /root/reference):  python -m oracle.make_golden
The fixtures are what travels to the GPU box; this script is the record of how they were made.
Inputs are regenerated from seeds by the tests (oracle.steps.make_real / make_rng), so only
outputs are stored: full loss trajectories, digests (sum / L2 / strided sample) of per-layer
activations and gradients at step 0, and of the weights, BN buffers and Adam moments at the end.
"""
import json
import os
import sys

import torch

from . import ref_harness, steps
from .trajectory import digest

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

DCGAN_CASES = {
    # name: (batch, steps, lr)
    "dcgan_b8_lr2e-4": (8, 100, 2e-4),       # trajectory parity (north_star: 100 steps)
    "dcgan_b8_lr1e-1": (8, 6, 0.1),          # the reference's default -mlr: saturates the BCE clamp
}
CGAN_CASES = {"cgan_b8_lr2e-4": (8, 12, 2e-4)}


def dcgan_inputs(batch, n_steps):
    real = steps.make_real(batch, n_steps=n_steps, seed=12345)
    rng = steps.make_rng(batch, n_steps=n_steps, seed=777)
    fixed = torch.randn(64, 100, 1, 1, generator=torch.Generator().manual_seed(4242))
    return real, rng, fixed


def cgan_inputs(batch, n_steps, n_classes=100):
    real = steps.make_real(batch, n_steps=n_steps, seed=12345)
    rng = steps.make_rng(batch, n_steps=n_steps, seed=777, dropout_dim=256)
    gen = torch.Generator().manual_seed(31337)
    labels = [steps.one_hot(torch.randint(0, n_classes, (batch,), generator=gen), n_classes)
              for _ in range(n_steps)]
    fixed_list = [torch.randn(10, 100, 1, 1, generator=gen) for _ in range(100)]
    fixed_labels = torch.vstack([steps.one_hot(torch.full((10,), i), n_classes) for i in range(100)])
    return real, labels, rng, fixed_list, fixed_labels


def _state_digest(sd):
    return {k: digest(v) for k, v in sd.items()}


def _opt_digest(osd):
    out = {}
    for idx, st in osd["state"].items():
        out[str(idx)] = {"step": float(st["step"]), "exp_avg": digest(st["exp_avg"]),
                         "exp_avg_sq": digest(st["exp_avg_sq"])}
    return out


def main():
    if not ref_harness.available():
        sys.exit("reference not present; golden fixtures can only be made in the build container")
    os.makedirs(OUT, exist_ok=True)
    meta = {"torch": torch.__version__, "threads": torch.get_num_threads(),
            "how": "python -m oracle.make_golden (drives /root/reference/train/*_trainer.py unmodified)"}
    for name, (b, n, lr) in DCGAN_CASES.items():
        real, rng, fixed = dcgan_inputs(b, n)
        ref = ref_harness.run_dcgan(real, rng, fixed, lr=lr, hook=True)
        hooks = {}
        for net in ("d", "g"):
            for layer, calls in ref["hooks"][net].items():
                # step 0: D is called A,B,C,D (4x), G once (+ once more by the eval branch)
                k = 4 if net == "d" else 1
                for ci, rec in enumerate(calls[:k]):
                    tag = f"{net}.{layer}.{'ABCD'[ci] if net == 'd' else 'fwd'}"
                    hooks[tag + ".out"] = digest(rec["out"])
                    if rec["grad"] is not None:
                        hooks[tag + ".grad"] = digest(rec["grad"])
        doc = {"meta": meta, "case": {"model": "DCGAN", "batch": b, "steps": n, "lr": lr},
               "losses_d": ref["losses_d"], "losses_g": ref["losses_g"], "step0": hooks,
               "d_state": _state_digest(ref["d_state"]), "g_state": _state_digest(ref["g_state"]),
               "opt_d": _opt_digest(ref["opt_d"]), "opt_g": _opt_digest(ref["opt_g"])}
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(doc, f)
        print(name, "loss_d[0..2]", ref["losses_d"][:3], "last", ref["losses_d"][-1])
    for name, (b, n, lr) in CGAN_CASES.items():
        real, labels, rng, fixed_list, fixed_labels = cgan_inputs(b, n)
        ref = ref_harness.run_cgan(real, labels, rng, fixed_list, lr=lr)
        doc = {"meta": meta, "case": {"model": "CGAN", "batch": b, "steps": n, "lr": lr},
               "losses_d": ref["losses_d"], "losses_g": ref["losses_g"],
               "d_state": _state_digest(ref["d_state"]), "g_state": _state_digest(ref["g_state"])}
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(doc, f)
        print(name, "loss_d[0..2]", ref["losses_d"][:3])


if __name__ == "__main__":
    main()
