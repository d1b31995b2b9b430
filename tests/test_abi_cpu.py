"""CPU-side checks of the C-ABI boundary: the library loads without a GPU or driver, exports every
symbol include/jck_b200.h declares, the ctypes binding covers all of them, and argument validation
answers with error codes (no compute is attempted here)."""
import ctypes
import os
import subprocess

import pytest

import __graft_entry__ as entry
from jck_generation_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    entry.build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    names = _lib.header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/jck_b200.h but not exported"
    assert sorted(_lib._SIGNATURES) == names, "ctypes binding and header disagree"


def test_library_does_not_link_the_driver():
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out, out


def test_bad_arguments_return_codes_not_crashes(lib):
    assert lib.jck_version() >= 100
    rc = lib.jck_conv_down(None, None, None, None, 0, 0, 0, 0, 0, 0, 0, 0, None)
    assert rc == -1
    assert b"conv_down" in lib.jck_last_error_string()
    assert lib.jck_adam(None, None, None, None, 0, 0.0, 0.0, 0.0, 0.0, None, None) == -1
    with pytest.raises(_lib.JckError):
        _lib.check(rc, "conv_down")


def test_sass_is_blackwell_native():
    """The tensor-core kernels must be tcgen05 / TMEM / TMA, not recompiled mma.sync."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not present")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")


def test_modules_refuse_cpu_tensors():
    import torch
    from jck_generation_b200.model import DCGAN
    d = DCGAN.Discriminator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.zeros(2, 3, 64, 64))
    g = DCGAN.Generator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        g(torch.zeros(2, 100, 1, 1))
