"""Per-kernel parity on the GPU, through the C ABI, against torch CPU operators."""
import pytest

from tests import kernel_checks as kc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _built():
    import __graft_entry__ as entry
    entry.build()


@pytest.mark.parametrize("spec", kc.all_cases(), ids=lambda s: "-".join(map(str, s)))
def test_kernel(spec):
    res = kc.run_case(*spec)
    for k, v in res.items():
        assert v <= kc.tolerance(spec[0], spec[2], k), f"{spec} {k}: {v}"


@pytest.mark.parametrize("n,d", [(3001, 100), (4096, 2048), (64, 100)])
def test_feature_moments(n, d):
    """FID feature moments on our kernels (column sums, hi/lo bf16 split, three tcgen05 GEMMs) against numpy's
    np.mean / np.cov (reference metrics.py:118-124) on correlated features with a large common offset."""
    import numpy as np
    import torch
    from jck_generation_b200 import ops
    g = torch.Generator().manual_seed(7)
    mix = torch.randn(d, d, generator=g) / d ** 0.5
    f = (torch.randn(n, d, generator=g) @ mix) * 2.0 + 5.0 + torch.randn(d, generator=g)
    mean, cov = ops.feature_moments(f.float().cuda())
    torch.cuda.synchronize()
    f64 = f.float().double().numpy()
    want_mean, want_cov = np.mean(f64, axis=0), np.cov(f64, rowvar=False)
    em = np.linalg.norm(mean.double().cpu().numpy() - want_mean) / np.linalg.norm(want_mean)
    ec = np.linalg.norm(cov.double().cpu().numpy() - want_cov) / np.linalg.norm(want_cov)
    assert em <= 1e-6 and ec <= 3e-5, (em, ec)
    assert float((cov - cov.t()).abs().max()) <= 1e-5 * float(cov.abs().max())
