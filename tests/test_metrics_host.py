"""Host logic of jck_generation_b200/metrics.py on the CPU, with the device pieces (feature extraction, moments, score kernel)
stubbed by numpy: superclass bookkeeping and the intra-FID of metrics.py:133-141 taken from already extracted features."""
import numpy as np
import torch

from jck_generation_b200.metrics import Metrics, SUPERCLASS


def _ref_fid(real, gen):
    """metrics.py:113-131 verbatim on numpy features"""
    from scipy.linalg import sqrtm
    mu1, sigma1 = np.mean(real, axis=0), np.cov(real, rowvar=False)
    mu2, sigma2 = np.mean(gen, axis=0), np.cov(gen, rowvar=False)
    diff = np.sum((mu1 - mu2) ** 2.0)
    covmean = sqrtm(sigma1.dot(sigma2))
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return diff + np.trace(sigma1 + sigma2 - 2.0 * covmean)


def _stub_metrics(real_feats, real_targets):
    m = object.__new__(Metrics)                       # skip the constructor: it needs a CUDA device by design
    m.device, m.feature, m.batch, m.comm = torch.device("cpu"), "logits", 128, None
    m.class_to_superclass = {c: s for s, cs in enumerate(SUPERCLASS) for c in cs}
    fake_targets = [i for i in range(100) for _ in range(10)]
    m.real_superclass_idx = {s: [i for i, t in enumerate(real_targets) if m.class_to_superclass[int(t)] == s] for s in range(20)}
    m.fake_superclass_idx = {s: [i for i, t in enumerate(fake_targets) if m.class_to_superclass[t] == s] for s in range(20)}
    m.real_features = real_feats
    m._moments = lambda f, sharded=False: (lambda a: (np.mean(a, axis=0), np.cov(a, rowvar=False)))(
        f.double().numpy() if torch.is_tensor(f) else np.asarray(f, dtype=np.float64))
    return m


def test_superclass_table_is_the_reference_partition():
    flat = sorted(c for cs in SUPERCLASS for c in cs)
    assert flat == list(range(100)) and all(len(cs) == 5 for cs in SUPERCLASS)
    assert SUPERCLASS[0] == [4, 30, 55, 72, 95] and SUPERCLASS[19] == [41, 69, 81, 85, 89]      # metrics.py:23-43


def test_intra_fid_from_features_matches_reference_formula():
    rng = np.random.default_rng(0)
    d = 6
    real_targets = rng.integers(0, 100, 4000)
    real = rng.normal(size=(4000, d)) + real_targets[:, None] * 0.01
    gen = torch.from_numpy(rng.normal(size=(1000, d)) * 1.3 + 0.2).float()
    m = _stub_metrics(real, real_targets)
    got = m._intra_fid_from(gen)
    want = 0.0
    for s in range(20):
        want += _ref_fid(real[m.real_superclass_idx[s]], gen.double().numpy()[m.fake_superclass_idx[s]])
    assert abs(got - want / 100) <= 1e-9 * abs(want)          # the reference divides the sum of 20 by 100 (metrics.py:141)
    m._extract = lambda images, real=False, generated=False: gen
    m._score = lambda logits, n, splits: 1.5
    score, fid, intra = m.evaluate_generated(torch.zeros(1000, 3, 4, 4), intra=True)
    assert score == 1.5 and abs(fid - _ref_fid(real, gen.double().numpy())) <= 1e-9 * abs(fid) and intra == got
    assert len(m.evaluate_generated(torch.zeros(1000, 3, 4, 4))) == 2
    # a generated set that is not the 1000-sample class-ordered one, or unknown real classes: NaN, not an exception
    assert np.isnan(m._intra_fid_from(gen[:64]))
    m.real_superclass_idx = {}
    assert np.isnan(m._intra_fid_from(gen))


REF_ROOT = "/root/reference"


def _load_reference_metrics():
    """the reference's own metrics.py, imported from where it lies (never copied); its module-level imports need the
    reference root on sys.path (`from utils import get_default_device`)"""
    import importlib.util
    import os
    import sys
    if not os.path.isfile(os.path.join(REF_ROOT, "metrics.py")):
        return None
    saved = {k: sys.modules.get(k) for k in ("utils", "metrics")}
    sys.path.insert(0, REF_ROOT)
    try:
        spec = importlib.util.spec_from_file_location("_jck_ref_metrics", os.path.join(REF_ROOT, "metrics.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(REF_ROOT)
        for k, v in saved.items():
            if v is None:
                m = sys.modules.get(k)
                if m is not None and (getattr(m, "__file__", "") or "").startswith(REF_ROOT):
                    del sys.modules[k]
            else:
                sys.modules[k] = v
    return mod


def test_formulas_against_the_live_reference_class():
    """Pin against the reference ITSELF (metrics.py:96-141), run here: its unmodified `inception_score`, `fid` and
    `intra_fid` methods on an instance whose private feature extractor is replaced by a table lookup (the constructor needs a
    checkpoint and CIFAR-100, metrics.py:51,56), against ours on the same features with the device pieces stubbed."""
    import pytest
    ref = _load_reference_metrics()
    if ref is None:
        pytest.skip("/root/reference is not present on this machine")
    rng = np.random.default_rng(1)
    d = 8
    real_targets = rng.integers(0, 100, 3000)
    real = (rng.normal(size=(3000, d)) + real_targets[:, None] * 0.02).astype(np.float32)
    logits = torch.from_numpy(rng.normal(size=(1000, d)).astype(np.float32) * 1.5 + 0.3)
    ours = _stub_metrics(real, real_targets)

    r = object.__new__(ref.Metrics)
    r.real_features = real
    r.real_superclass_idx, r.fake_superclass_idx = ours.real_superclass_idx, ours.fake_superclass_idx

    def extract(images, real=False, softmax=False):            # images: DataLoader over ROW INDICES into `logits`
        rows = torch.cat([b for b in images]).long()
        f = logits[rows]
        return torch.softmax(f, dim=1).numpy() if softmax else f.numpy()
    r._Metrics__extract_features = extract
    idx = torch.arange(1000)
    loader = torch.utils.data.DataLoader(idx, batch_size=128)
    want_is = r.inception_score(loader, splits=10)
    want_fid = r.fid(loader)
    want_intra = r.intra_fid(idx)

    from tests import incep_emul as emu
    scores = torch.zeros(10)
    emu.inception_score(logits, 10, scores)                      # the score kernel's restatement (GPU test: kernel == scipy)
    assert abs(float(scores.mean()) - want_is) <= 1e-5 * want_is
    f64 = logits
    assert abs(ours._fid_from(f64) - want_fid) <= 1e-6 * abs(want_fid)
    assert abs(ours._intra_fid_from(f64) - want_intra) <= 1e-6 * abs(want_intra)
