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
