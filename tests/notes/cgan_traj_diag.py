"""Diagnostic (GPU): per-step deviation of the CGAN trajectory, free-running vs teacher-forced."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import parity
from oracle import make_golden

gold = json.load(open(os.path.join(os.path.dirname(__file__), "..", "golden", "cgan_b8_lr2e-4.json")))
n = gold["case"]["steps"]
real, labels, rng, _, _ = make_golden.cgan_inputs(8, n)
for tf in (True, False):
    got, want, _ = parity.cgan_trajectory(torch.float32, batch=8, steps=n, lr=2e-4, real=real, labels=labels, rng=rng, teacher_forced=tf)
    print("teacher_forced" if tf else "free-running")
    for i in range(n):
        print(i, " ".join(f"{k}:{got[i][k]:.5f}/{want[i][k]:.5f}" for k in ("loss_d", "loss_g", "gp", "err_real", "err_fake", "x_d")),
              "gold", f"{gold['losses_d'][i]:.5f} {gold['losses_g'][i]:.5f}")
