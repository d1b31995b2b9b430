"""CPU checks of the Inception-v3 host graph (jck_generation_b200/inception.py) with the kernel primitives emulated in torch
(tests/incep_emul.py): buffer borders, tap shifts, stride-2 patch matrices, concat offsets and BatchNorm folding against
torchvision's inception_v3 -- the network the reference evaluates with (metrics.py:46-52,87)."""
import pytest
import torch

from tests import incep_emul as emu
from tests.incep_fixture import calibrated_inception


@pytest.fixture(scope="module")
def model():
    return calibrated_inception(seed=1)


def test_graph_matches_torchvision_fp32(model):
    from jck_generation_b200.inception import InceptionV3
    x = torch.randn(1, 3, 299, 299, generator=torch.Generator().manual_seed(0))
    acts = {}
    hooks = [getattr(model, n).register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, o))
             for n in ("Conv2d_2b_3x3", "Mixed_5d", "Mixed_6a", "Mixed_6e", "Mixed_7a", "Mixed_7c")]
    with torch.no_grad():
        ref = model(x)
    for h in hooks:
        h.remove()
    net = InceptionV3(model.state_dict(), device="cpu", K=emu, act_dtype=torch.float32)
    out = net.forward(x)
    keys = {"Conv2d_2b_3x3": "c2b"}
    for n, r in acts.items():
        mine = net._bufs[(keys.get(n, n + ".out"), 1)].interior().permute(0, 3, 1, 2)
        assert float((mine - r).norm() / r.norm()) < 5e-4, n
    assert float((out - ref).norm() / ref.norm()) < 1e-4
    assert net.launches == 94 + 5 + 13 + 1 + 1            # convs, patch matrices (4 stride-2 + the fused stem), pools, avgpool, fc
    # borders were never written
    for b in net._bufs.values():
        if not isinstance(b, torch.Tensor) and (b.py or b.px):
            full = b.t.view(b.B, b.Hb, b.Wb, b.ld)
            assert float(full[:, :b.py].abs().sum() + full[:, :, :b.px].abs().sum()) == 0.0


def test_split_precision_graph_free_running_matches_torchvision(model):
    """precision="split": activations and weights as hi + lo bf16 pairs, products hi*hi + hi*lo + lo*hi (the tap list and
    the weight blocks tripled by the host graph, jck_conv_gemm itself unchanged).  FREE RUNNING against torchvision fp32:
    5e-4 at the logits, <= 3e-3 at every block -- where plain bf16 storage is 17 % off at the logits of this (random-weight,
    chaotic) test network.  This is the mode that pins the extractor against the reference's arithmetic end to end."""
    from jck_generation_b200.inception import InceptionV3
    x = torch.randn(1, 3, 299, 299, generator=torch.Generator().manual_seed(0))
    acts = {}
    hooks = [getattr(model, n).register_forward_hook(lambda m, i, o, n=n: acts.__setitem__(n, o))
             for n in ("Conv2d_2b_3x3", "Mixed_5d", "Mixed_6e", "Mixed_7c")]
    with torch.no_grad():
        ref = model(x)
    for h in hooks:
        h.remove()
    net = InceptionV3(model.state_dict(), device="cpu", K=emu, precision="split")
    out = net.forward(x)
    for n, r in acts.items():
        mine = net._bufs[({"Conv2d_2b_3x3": "c2b"}.get(n, n + ".out"), 1)].value().permute(0, 3, 1, 2)
        assert float((mine - r).norm() / r.norm()) < 5e-3, n
    assert float((out - ref).norm() / ref.norm()) < 1.5e-3
    assert net.launches == 94 + 2 * 4 + 1 + 13 + 1 + 1      # each stride-2 patch matrix once per plane
    plain = InceptionV3(model.state_dict(), device="cpu", K=emu).forward(x)
    assert float((plain - ref).norm() / ref.norm()) > 20 * float((out - ref).norm() / ref.norm())
    # pool3 features and the generated-image entry (stem fused with resize / normalise) in split mode
    p3 = InceptionV3(model.state_dict(), feature="pool3", device="cpu", K=emu, precision="split")
    fake = torch.tanh(torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)))
    f = p3.forward_generated(fake)
    pre = torch.nn.functional.interpolate(0.5 * fake + 0.5, size=(299, 299), mode="bilinear", align_corners=False)
    pre = (pre - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    feats = {}
    h = model.avgpool.register_forward_hook(lambda m, i, o: feats.__setitem__("p", o.flatten(1)))
    with torch.no_grad():
        model(pre)
    h.remove()
    assert float((f - feats["p"]).norm() / feats["p"].norm()) < 3e-3


def test_pool3_features_and_generated_entry(model):
    from jck_generation_b200.inception import InceptionV3
    net = InceptionV3(model.state_dict(), feature="pool3", device="cpu", K=emu, act_dtype=torch.float32)
    fake = torch.tanh(torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)))
    f = net.forward_generated(fake)
    # the reference's eval branch (dcgan_trainer.py:203-207) followed by the trunk up to avgpool
    pre = torch.nn.functional.interpolate(0.5 * fake + 0.5, size=(299, 299), mode="bilinear", align_corners=False)
    pre = (pre - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    feats = {}
    h = model.avgpool.register_forward_hook(lambda m, i, o: feats.__setitem__("p", o.flatten(1)))
    with torch.no_grad():
        model(pre)
    h.remove()
    assert f.shape == (1, 2048)
    assert float((f - feats["p"]).norm() / feats["p"].norm()) < 1e-4


def test_inception_score_emulator_matches_scipy():
    import numpy as np
    from scipy.stats import entropy
    logits = torch.randn(200, 100, generator=torch.Generator().manual_seed(4)) * 2
    scores = torch.zeros(4)
    emu.inception_score(logits, 4, scores)
    preds = torch.softmax(logits, 1).numpy()
    for k in range(4):
        part = preds[k * 50:(k + 1) * 50]
        py = part.mean(0)
        want = np.exp(np.mean([entropy(part[i], py) for i in range(50)]))
        assert abs(float(scores[k]) - want) < 1e-4 * want


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour WITHOUT a CUDA device")
def test_no_cpu_path_in_the_product(model, tmp_path, monkeypatch):
    """the product has no CPU route: Metrics refuses to construct, and the extractor with the real backend refuses CPU tensors"""
    from jck_generation_b200.inception import InceptionV3
    from jck_generation_b200.metrics import Metrics
    monkeypatch.chdir(tmp_path)
    with pytest.raises(RuntimeError, match="no CPU path"):
        Metrics(None)
    net = InceptionV3(model.state_dict(), device="cpu", use_graph=False)          # K defaults to the C ABI
    with pytest.raises(AssertionError, match="CUDA"):
        net.forward(torch.zeros(1, 3, 299, 299))
    with pytest.raises(AssertionError, match="bf16"):
        InceptionV3(model.state_dict(), device="cpu", act_dtype=torch.float32)   # fp32 buffers exist only for the emulator
