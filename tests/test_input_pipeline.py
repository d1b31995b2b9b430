"""Input pipeline: the oracle's restatement of Pillow's resample + torchvision's ToTensor / Normalize is
pinned against Pillow and torchvision themselves (CPU), the product's coefficient tables against the oracle's (CPU), and the
CUDA kernels against the oracle (GPU) -- all bit for bit, as the reference's transform (dcgan_data_preprocessor.py:38-49,
cgan_data_preprocessor.py:11-16,51-62) is byte / fixed-point arithmetic followed by two IEEE fp32 operations."""
import numpy as np
import pytest
import torch

from oracle import pil_resize as pr

CASES = [(32, 32, 64, 64), (32, 32, 299, 299), (28, 28, 64, 64), (48, 40, 64, 53), (64, 64, 32, 32), (32, 32, 32, 32)]
HALF = ([0.5] * 3, [0.5] * 3)
IMNET = ([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])


@pytest.mark.parametrize("h,w,oh,ow", CASES)
def test_oracle_resize_equals_pillow(h, w, oh, ow):
    from PIL import Image
    rng = np.random.default_rng(h * 1000 + ow)
    for img in (rng.integers(0, 256, (h, w, 3), dtype=np.uint8), np.full((h, w, 3), 255, np.uint8), np.zeros((h, w, 3), np.uint8)):
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(pr.resize_u8(img, oh, ow), want)


@pytest.mark.parametrize("size,oh,ow,norm", [(64, 64, 64, HALF), ((299, 299), 299, 299, IMNET)])
def test_oracle_transform_equals_torchvision(size, oh, ow, norm):
    """the reference's two Compose pipelines, verbatim"""
    from PIL import Image
    import torchvision.transforms as tt
    rng = np.random.default_rng(3)
    t = tt.Compose([tt.Resize(size), tt.ToTensor(), tt.Normalize(mean=norm[0], std=norm[1])])
    for _ in range(3):
        img = rng.integers(0, 256, (32, 32, 3), dtype=np.uint8)
        assert np.array_equal(pr.transform(img, oh, ow, *norm), t(Image.fromarray(img)).numpy())


def test_one_hot_equals_reference_encoder():
    labels = [3, 0, 99, 42]
    ref = torch.stack([torch.LongTensor([1 if i == l else 0 for i in range(100)]) for l in labels])   # cgan_data_preprocessor.py:15-16
    assert np.array_equal(pr.one_hot(labels, 100), ref.numpy())


@pytest.mark.parametrize("a,b", [(32, 64), (32, 299), (28, 64), (40, 53), (64, 32), (299, 64), (33, 77)])
def test_product_tables_equal_oracle(a, b):
    from jck_generation_b200.preprocess.device_pipeline import bilinear_tables
    b1, k1 = pr.bilinear_coeffs(a, b)
    b2, k2 = bilinear_tables(a, b)
    assert np.array_equal(b1, b2) and np.array_equal(k1, k2)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,size,norm", [(32, 32, 64, HALF), (32, 32, (299, 299), IMNET), (28, 28, 64, HALF), (48, 40, (64, 53), HALF),
                                           (64, 64, (32, 32), HALF), (64, 64, 64, HALF)])
def test_device_pipeline_bit_exact(h, w, size, norm):
    import __graft_entry__ as entry
    entry.build()
    from jck_generation_b200.preprocess.device_pipeline import DeviceImageLoader
    rng = np.random.default_rng(5)
    n = 37
    data = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    data[0], data[1] = 0, 255                                    # saturated images
    targets = rng.integers(0, 100, n).tolist()
    loader = DeviceImageLoader(data, targets, 16, size, norm[0], norm[1], shuffle=True, n_classes=100)
    torch.manual_seed(77)
    ref_idx = [b for b in torch.utils.data.DataLoader(torch.arange(n), 16, shuffle=True)]   # the reference loader's batches
    assert len(loader) == 3 and len(loader.dataset) == n
    torch.manual_seed(77)                                        # same global RNG state -> same permutation as the reference
    for (x, y), idx in zip(loader, ref_idx):
        idx = idx.tolist()
        want = np.stack([pr.transform(data[i], loader.Ho, loader.Wo, *norm) for i in idx])
        assert x.is_cuda and x.dtype == torch.float32 and tuple(x.shape) == want.shape
        assert np.array_equal(x.cpu().numpy(), want)
        assert np.array_equal(y.cpu().numpy(), pr.one_hot([targets[i] for i in idx], 100))
    plain = DeviceImageLoader(data, targets, 16, size, norm[0], norm[1], shuffle=False)
    x, y = next(iter(plain))
    assert y.dtype == torch.int64 and y.tolist() == targets[:16]


@pytest.mark.gpu
def test_preprocessors_use_device_pipeline(tmp_path, monkeypatch):
    from types import SimpleNamespace
    monkeypatch.chdir(tmp_path)
    from jck_generation_b200.preprocess.cgan_data_preprocessor import CGANDataPreprocessor
    from jck_generation_b200.preprocess.dcgan_data_preprocessor import DCGANDataPreprocessor
    args = SimpleNamespace(batch_size=32, num_worker=0, synthetic=1, synthetic_u8=1, synthetic_batches=3, log_level="error",
                           model_name="t", n_classes=100)
    for cls, onehot in ((DCGANDataPreprocessor, False), (CGANDataPreprocessor, True)):
        pre = cls(args)
        pre.transform_data()
        train, metric = pre.get_data_loader()
        x, y = next(iter(train))
        assert x.is_cuda and tuple(x.shape) == (32, 3, 64, 64) and float(x.min()) >= -1.0 and float(x.max()) <= 1.0
        assert tuple(y.shape) == ((32, 100) if onehot else (32,))
        xm, _ = next(iter(metric))
        assert tuple(xm.shape)[1:] == (3, 299, 299) and len(metric.dataset) == 96


@pytest.mark.gpu
def test_trainer_with_device_pipeline_and_metrics(tmp_path, monkeypatch):
    """main.py:83-96 end to end on the device: uint8 dataset in HBM -> DeviceImageLoader -> DCGANTrainer.train() (CUDA graph)
    with the eval branch live (Metrics on the jck Inception extractor, real features from the device metric loader)."""
    import argparse
    monkeypatch.chdir(tmp_path)
    # no Inception checkpoint offline: Metrics raises like the reference unless random weights are asked for explicitly
    monkeypatch.setenv("JCK_METRICS_RANDOM_WEIGHTS", "1")
    from jck_generation_b200.model import DCGAN
    from jck_generation_b200.preprocess.dcgan_data_preprocessor import DCGANDataPreprocessor
    from jck_generation_b200.train.dcgan_trainer import DCGANTrainer
    args = argparse.Namespace(epoch=1, max_learning_rate=2e-4, model_path="t", log_file=0, batch_size=32, num_worker=0,
                              synthetic=1, synthetic_u8=1, synthetic_batches=4, dtype="bf16", cuda_graph=1, metrics=1,
                              save_path=str(tmp_path))
    data = DCGANDataPreprocessor(args)
    data.transform_data()
    tr = DCGANTrainer(args, DCGAN.Generator(), DCGAN.Discriminator(), data)
    assert tr.metric is not None and tr.metric.real_features.shape == (128, 100)
    losses_d, losses_g = tr.train()
    assert len(losses_d) == 4 and all(v == v and abs(v) < 200 for v in losses_d + losses_g)
