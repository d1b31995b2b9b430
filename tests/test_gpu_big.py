"""Parity of the train step AT THE BENCHMARKED SIZES (BASELINE.json configs: DCGAN batch 128 and 512, CGAN batch 256)
against the CPU oracle, through the public step API and the C ABI.

Why these exist: at a batch of 8 every tcgen05 kernel runs at most one tile per CTA.  At 512 images per GPU (x3 BatchNorm
groups in the discriminator's [real | fake | x_hat] pass) each CTA pair walks 5-20 tiles: TMEM double buffering, ring phase
carry-over across tiles, register-resident BatchNorm partial sums flushed on a group change and the split-K ranges of the
weight gradients are only exercised here (and in the per-kernel cases of tests/kernel_checks.py:big_cases).

Tolerances are north_star's where the arithmetic allows them and otherwise a MEASURED, explained replacement
(numbers: tests/notes/big_parity.py on a B200, profiles/r02/big_parity.json):

* bf16 activations: <= 1e-2 for every layer except the discriminator's conv3 / conv4 on GENERATED images (groups B and D),
  which sit 8 and 9 bf16 layers deep (five generator layers, then D) and measure 1.00e-2 / 1.19e-2: bound 1.3e-2.
* bf16 losses and D outputs: <= 5e-3 (measured <= 1.5e-3).
* bf16 gradients: every parameter gradient is held to the error torch's OWN bf16 autocast makes on the same tensor
  (tests/parity.py:autocast_envelope, same weights / inputs / batch) x 1.1 + 5e-3.  Measured at batch 128: ours is BELOW
  autocast's error on all 26 tensors (e.g. D.conv1.weight 8.3 % vs 8.6 %, G.conv1.weight 16.3 % vs 16.6 %).  1e-2 is not
  reachable by bf16 operands on this network: BatchNorm backward removes the components of the incoming gradient along
  (1, x_hat), which at N(0, .02) initialisation is most of it, so 2^-9 operand rounding is amplified 10-50x.  Thirty
  optimiser steps off the initialisation the same comparison gives 2-4 % (D) and 1.5-5 % (G; bounds 6 % / 8 %) -- except G.conv1.weight,
  whose fp32 gradient is dominated by the sampling noise of z (13 % for us, 14 % for autocast).
* fp32 mode at batch 128: activations and scalars <= 1e-4 (measured 2.4e-6 / 1.8e-5); gradients <= 5e-3 -- at this size
  ~25 of the 32 M LeakyReLU / ReLU pre-activations lie within fp32 summation-order noise of zero (|pre| < 3e-6) and take the
  other branch than the oracle's (tests/parity.py:_kink_flips); each flip moves a gradient tensor by ~1e-4..1e-3.  The
  flipped elements are asserted to be rounding-small; the batch-8 tests (test_gpu_step.py) hold a flip-free draw to 1e-4.
"""
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu

DEEP_FAKE = ("d_act.B.conv3", "d_act.B.conv4", "d_act.D.conv3", "d_act.D.conv4")
SCALARS = ("scalar.loss_d", "scalar.loss_g", "scalar.x_d", "scalar.z1_gd", "scalar.z2_gd", "scalar.err_real",
           "scalar.err_fake", "scalar.gp")


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


def _check_bf16_forward(errs, scalar_tol=5e-3):
    for k, v in errs.items():
        if k.startswith(("d_act", "g_act", "fake_raw")):
            assert v <= (1.3e-2 if k in DEEP_FAKE else 1e-2), f"{k}: {v}"
    for k in SCALARS:
        assert errs[k] <= scalar_tol, f"{k}: {errs[k]}"


def _check_bf16_grads(errs, env):
    for k, e in env.items():
        assert errs[k] <= 1.1 * e + 5e-3, f"{k}: ours {errs[k]:.4f} vs torch bf16 autocast {e:.4f}"


@pytest.mark.parametrize("batch", [128, 512])
def test_dcgan_bf16_step_at_benchmark_batch(batch):
    """BASELINE configs[0] (batch 128) and configs[2] (batch 512), tcgen05 mode: every activation of the four discriminator
    passes and the generator, every logged scalar, every parameter gradient."""
    errs = parity.dcgan_step_parity(torch.bfloat16, batch=batch)
    _check_bf16_forward(errs)
    _check_bf16_grads(errs, parity.autocast_envelope(batch))
    assert errs["gp_grads"] <= 0.15, errs["gp_grads"]            # the penalty's input gradient: measured 0.10


def test_dcgan_bf16_gradients_off_initialisation():
    """The same step 30 optimiser steps away from the N(0, .02) initialisation (the oracle trains, the CUDA side starts
    from its state): the BatchNorm-backward amplification of bf16 rounding shrinks -- D's gradients 2-4 %, G's 1.5-5 %
    except conv1.weight -- and stays within torch autocast's own error."""
    errs = parity.dcgan_step_parity(torch.bfloat16, batch=128, warm_steps=30)
    # scalars at north_star's bf16 tolerance here: thirty steps in, D(G(z)) after the update is a small mean (its relative error
    # measured 4.9e-3 and 5.8e-3 in two runs of the same code -- the statistics' atomics commute only up to fp32 rounding)
    _check_bf16_forward(errs, scalar_tol=1e-2)
    env = parity.autocast_envelope(128, warm_steps=30)
    _check_bf16_grads(errs, env)
    for k, v in errs.items():
        if k.startswith("d_grad"):
            assert v <= 6e-2, f"{k}: {v}"
        elif k.startswith("g_grad") and k != "g_grad.conv1.weight":
            assert v <= 8e-2, f"{k}: {v}"


def test_dcgan_fp32_step_at_batch_128():
    errs = parity.dcgan_step_parity(torch.float32, batch=128)
    for k, v in errs.items():
        if k.startswith(("d_act", "g_act", "fake_raw")) or k in SCALARS:
            assert v <= 1e-4, f"{k}: {v}"
        elif k.startswith(("d_grad", "g_grad", "gp_grads")):
            assert v <= 5e-3, f"{k}: {v}"
    assert errs["kink.flips"] <= 200 and errs["kink.worst_pre"] < 1e-4, (errs["kink.flips"], errs["kink.worst_pre"])


@pytest.mark.parametrize("nc,n_classes", [(3, 100), (1, 10)])
def test_cgan_bf16_step_at_batch_256(nc, n_classes):
    """BASELINE configs[1] (CGAN, batch 256) at the reference's native shape and at the config's 1-channel / 10-class shape.
    D's gradients contain the second-order terms of the back-propagated penalty.  Measured: scalars <= 5e-4, D gradients
    <= 6 %, G gradients <= 12 % except conv1.weight 15.7 % (cf. the DCGAN envelope above), penalty input gradient 10 %."""
    e = parity.cgan_step_parity(torch.bfloat16, batch=256, nc=nc, n_classes=n_classes)
    assert e["fake_raw"] <= 1e-2, e["fake_raw"]
    for k in SCALARS:
        assert e[k] <= 5e-3, f"{k}: {e[k]}"
    for k, v in e.items():
        if k.startswith("d_grad"):
            assert v <= 0.1, f"{k}: {v}"
        elif k.startswith("g_grad"):
            assert v <= 0.2, f"{k}: {v}"
    assert e["gp_grads"] <= 0.15, e["gp_grads"]


def test_cgan_fp32_step_at_batch_64():
    e = parity.cgan_step_parity(torch.float32, batch=64)
    for k in SCALARS:
        assert e[k] <= 1e-4, f"{k}: {e[k]}"
    assert e["fake_raw"] <= 1e-4
    for k, v in e.items():
        if k.startswith(("d_grad", "g_grad", "gp_grads")):
            assert v <= 5e-3, f"{k}: {v}"
    assert e["kink.worst_pre"] < 1e-4
