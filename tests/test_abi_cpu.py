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


def test_metrics_and_input_entry_points_validate_arguments(lib):
    """error behaviour of the FID / IS and input-pipeline entry points: codes and messages, before any CUDA call"""
    I = ctypes.c_int
    buf = ctypes.create_string_buffer(4096)
    p = ctypes.cast(buf, ctypes.c_void_p)

    def geom(*v):
        return ctypes.cast((I * len(v))(*v), ctypes.c_void_p)

    def err():
        return lib.jck_last_error_string().decode()
    # jck_conv_gemm: null pointers, short geometry, too many taps, unaligned pitch, row space not B x Hq x Wq
    assert lib.jck_conv_gemm(None, 64, None, None, None, None, 64, None, 0, None) == -1 and "conv_gemm" in err()
    g19 = [128, 64, 64, 1, 8, 16, 0, 0, 8, 16, 8, 16, 0, 0, 0, 1, 1, 128, 0]
    assert lib.jck_conv_gemm(p, 64, p, None, None, p, 64, geom(*g19[:10]), 10, None) == -1
    bad = list(g19); bad[3] = 100                   # more than the 96 taps (5 x 5 taps x 3 partial products of the split mode)
    assert lib.jck_conv_gemm(p, 64, p, None, None, p, 64, geom(*(bad + [0] * 100)), 119, None) == -1 and "geometry" in err()
    # split-precision output asks for bf16 and an aligned plane offset
    assert lib.jck_conv_gemm(p, 64, p, None, None, p, 63, geom(*(g19 + [3])), 20, None) == -1 and "split output" in err()
    assert lib.jck_pool3_split(p, geom(8, 8, 0, 0, 0), 64, 4, p, geom(8, 8, 0, 0, 0), 64, 0, 1, 8, 8, 64, 1, 1, 8, 8, 1, None) == -1
    assert lib.jck_conv_gemm(p, 60, p, None, None, p, 64, geom(*g19), 19, None) == -2 and "aligned" in err()
    bad = list(g19); bad[0] = 100
    assert lib.jck_conv_gemm(p, 64, p, None, None, p, 64, geom(*bad), 19, None) == -1 and "row space" in err()
    # jck_im2col: patch pitch too small / not a multiple of 8
    g5 = geom(8, 8, 0, 0, 0)
    assert lib.jck_im2col(p, g5, 8, p, 1, 8, 8, 8, 3, 3, 2, 2, 0, 0, 3, 3, 64, None) == -1 and "Kp" in err()
    assert lib.jck_im2col(p, g5, 8, p, 1, 8, 8, 8, 3, 3, 2, 2, 0, 0, 3, 3, 76, None) == -1
    # jck_pool3: channels not a multiple of 8; a max window that leaves the image
    assert lib.jck_pool3(p, g5, 12, p, g5, 12, 1, 8, 8, 12, 2, 0, 3, 3, 0, None) == -2 and "multiples of 8" in err()
    assert lib.jck_pool3(p, g5, 8, p, g5, 8, 1, 8, 8, 8, 2, 0, 4, 4, 0, None) == -1 and "window" in err()
    assert lib.jck_pool3(p, g5, 8, p, g5, 8, 1, 8, 8, 8, 2, 0, 3, 3, 5, None) == -1
    # the rest: null / out-of-range arguments
    three = ctypes.cast((ctypes.c_float * 3)(0, 0, 0), ctypes.c_void_p)
    assert lib.jck_global_avgpool(None, None, None, 1, 1, 1, None) == -1
    assert lib.jck_resize_norm(p, p, 1, 4, 8, 8, 8, 8, 4, 1.0, 0.0, three, three, None) == -1 and "C <= 3" in err()
    assert lib.jck_stem_patches(p, p, 1, 8, 8, 2, 2, 1.0, 0.0, three, three, None) == -1
    assert lib.jck_inception_score(p, 0, 100, 10, p, None) == -1
    assert lib.jck_u8_resize_norm(p, None, p, 1, 8, 8, 5, 16, 16, p, p, 3, p, p, 3, three, three, None) == -1 and "C <= 4" in err()
    assert lib.jck_u8_resize_norm(p, None, p, 1, 8, 8, 3, 16, 16, None, None, 0, p, p, 3, three, three, None) == -1 and "tables" in err()
    assert lib.jck_one_hot_i64(None, None, None, 1, 10, None) == -1
