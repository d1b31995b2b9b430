"""Pin the oracle (CPU restatement) against the reference's own outputs.

Two pins:
  * tests/golden/*.json -- digests frozen from the unmodified reference step loops by
    oracle/make_golden.py (runs everywhere, incl. the GPU box where /root/reference is absent);
  * the live reference, bit for bit, where /root/reference exists (build container only).
"""
import json
import os

import pytest
import torch

from oracle import make_golden, ref_harness, steps, trajectory
from oracle.trajectory import digest, digest_close

# golden digests were produced on the build container's CPU; another host may pick other oneDNN
# kernels, so cross-host comparisons are tolerance-based.  Same-host comparisons are bit-exact.
RTOL_SAME_BUILD = 5e-4


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name + ".json")) as f:
        return json.load(f)


def _check_states(gold, got, what, rtol):
    for k, dg in gold.items():
        if k.endswith("num_batches_tracked"):
            assert dg["sum"] == float(got[k].sum()), k
            continue
        ok, why = digest_close(dg, digest(got[k]), rtol=rtol, atol=1e-7)
        assert ok, f"{what}.{k}: {why}"


def test_dcgan_oracle_matches_golden_first_steps(golden_dir):
    """8 steps of the 100-step golden trajectory + per-layer digests at step 0."""
    gold = _load(golden_dir, "dcgan_b8_lr2e-4")
    c = gold["case"]
    n = 8
    real, rng, fixed = make_golden.dcgan_inputs(c["batch"], c["steps"])
    out = trajectory.run_dcgan(real[:n], rng[:n], fixed, lr=c["lr"], capture_first=True, emulate_eval=False)
    for i in range(n):
        assert out["losses_d"][i] == pytest.approx(gold["losses_d"][i], rel=RTOL_SAME_BUILD), i
        assert out["losses_g"][i] == pytest.approx(gold["losses_g"][i], rel=RTOL_SAME_BUILD), i
    cap = out["first"]["capture"]
    for tag, dg in gold["step0"].items():
        net, layer, which, kind = tag.split(".")
        if net == "d":
            src = cap["d_acts"] if kind == "out" else cap["d_act_grads"]
            t = src.get(which, {}).get(layer)
        else:
            src = cap["g_acts"] if kind == "out" else cap["g_act_grads"]
            t = src.get(layer)
        if t is None:      # conv5 of D is inside the head in the oracle's tap list
            continue
        ok, why = digest_close(dg, digest(t), rtol=RTOL_SAME_BUILD, atol=1e-9)
        assert ok, f"{tag}: {why}"


@pytest.mark.slow
def test_dcgan_oracle_matches_golden_full_trajectory(golden_dir):
    gold = _load(golden_dir, "dcgan_b8_lr2e-4")
    c = gold["case"]
    real, rng, fixed = make_golden.dcgan_inputs(c["batch"], c["steps"])
    out = trajectory.run_dcgan(real, rng, fixed, lr=c["lr"])
    # GAN trajectories amplify rounding differences between hosts; same build => tight
    for i in range(c["steps"]):
        assert out["losses_d"][i] == pytest.approx(gold["losses_d"][i], rel=2e-2, abs=2e-2), i
        assert out["losses_g"][i] == pytest.approx(gold["losses_g"][i], rel=2e-2, abs=2e-2), i
    _check_states(gold["d_state"], out["d_state"], "d_state", rtol=2e-2)
    _check_states(gold["g_state"], out["g_state"], "g_state", rtol=2e-2)


def test_dcgan_default_lr_saturates_like_reference(golden_dir):
    """-mlr default 0.1 (main.py:54) drives D into the BCE log clamp: loss_d == 110 exactly
    (100 + 0 + 10*1: err_real = -0.9*(-100)... see SURVEY 5.6)."""
    gold = _load(golden_dir, "dcgan_b8_lr1e-1")
    c = gold["case"]
    real, rng, fixed = make_golden.dcgan_inputs(c["batch"], c["steps"])
    out = trajectory.run_dcgan(real, rng, fixed, lr=c["lr"])
    assert gold["losses_d"][1] == 110.0
    for i in range(c["steps"]):
        assert out["losses_d"][i] == pytest.approx(gold["losses_d"][i], rel=1e-3), i
        assert out["losses_g"][i] == pytest.approx(gold["losses_g"][i], rel=1e-3), i


def test_cgan_oracle_matches_golden(golden_dir):
    gold = _load(golden_dir, "cgan_b8_lr2e-4")
    c = gold["case"]
    n = 4
    real, labels, rng, fixed_list, fixed_labels = make_golden.cgan_inputs(c["batch"], c["steps"])
    out = trajectory.run_cgan(real[:n], labels[:n], rng[:n], torch.vstack(fixed_list), fixed_labels,
                              lr=c["lr"], emulate_eval=False)
    for i in range(n):
        assert out["losses_d"][i] == pytest.approx(gold["losses_d"][i], rel=RTOL_SAME_BUILD), i
        assert out["losses_g"][i] == pytest.approx(gold["losses_g"][i], rel=RTOL_SAME_BUILD), i


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference only exists in the build container")
def test_oracle_bit_exact_vs_live_reference():
    b, n = 4, 2
    real = steps.make_real(b, n_steps=n)
    rng = steps.make_rng(b, n_steps=n, seed=3)
    fixed = torch.randn(64, 100, 1, 1, generator=torch.Generator().manual_seed(1))
    ref = ref_harness.run_dcgan(real, rng, fixed, lr=2e-4)
    orc = trajectory.run_dcgan(real, rng, fixed, lr=2e-4)
    assert ref["losses_d"] == orc["losses_d"] and ref["losses_g"] == orc["losses_g"]
    for k in ref["d_state"]:
        assert torch.equal(ref["d_state"][k], orc["d_state"][k]), k
    for k in ref["g_state"]:
        assert torch.equal(ref["g_state"][k], orc["g_state"][k]), k


@pytest.mark.skipif(not ref_harness.available(), reason="/root/reference only exists in the build container")
def test_oracle_modules_match_reference_state_dict():
    from oracle import models
    g_ref, d_ref = ref_harness.reference_modules(seed=12345)
    g, d = models.build("DCGAN", seed=12345)
    assert list(g.state_dict()) == list(g_ref.state_dict())
    assert list(d.state_dict()) == list(d_ref.state_dict())
    for k, v in g_ref.state_dict().items():
        assert torch.equal(v, g.state_dict()[k]), k
    for k, v in d_ref.state_dict().items():
        assert torch.equal(v, d.state_dict()[k]), k
    g.load_state_dict(g_ref.state_dict(), strict=True)
    d_ref.load_state_dict(d.state_dict(), strict=True)
