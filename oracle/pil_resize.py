"""TEST INFRASTRUCTURE (CPU oracle): numpy restatement of the reference's input transform

    tt.Compose([tt.Resize(size), tt.ToTensor(), tt.Normalize(mean, std)])       preprocess/dcgan_data_preprocessor.py:38-49
    OneHotEncoder(label_count)                                                  preprocess/cgan_data_preprocessor.py:11-16

for 8-bit RGB images.  The arithmetic lives in third-party code absent from /root/reference: Pillow (here 12.2.0,
src/libImaging/Resample.c: `precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`,
`ImagingResampleVertical_8bpc` -- two separable passes, 22-bit fixed-point coefficients, rounding to uint8 after EACH pass)
and torchvision (ToTensor = uint8 / 255 in fp32, Normalize = (x - mean) / std in fp32).  Pinned in
tests/test_input_pipeline.py against Pillow + torchvision themselves, bit for bit.
Only tests/ may import this module."""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def bilinear_coeffs(in_size, out_size):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the triangle filter (support 1.0) over the whole axis:
    -> (bounds int32 [out, 2] = (xmin, count), coefficients int32 [out, ksize])"""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)          # C (int) cast truncates toward zero; the argument is > -1 here
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            t = -t if t < 0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        for x in range(ksize):
            v = w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _clip8(acc):
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_u8(img, out_h, out_w):
    """img uint8 [H, W, C] -> uint8 [out_h, out_w, C]: horizontal pass, then vertical pass on the rounded result
    (ImagingResampleInner; a pass is skipped when that axis keeps its size)."""
    H, W, C = img.shape
    cur = img
    if out_w != W:
        bounds, kk = bilinear_coeffs(W, out_w)
        tmp = np.zeros((H, out_w, C), dtype=np.uint8)
        for xx in range(out_w):
            x0, n = bounds[xx]
            acc = np.full((H, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for x in range(n):
                acc += cur[:, x0 + x, :].astype(np.int64) * int(kk[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        cur = tmp
    if out_h != H:
        bounds, kk = bilinear_coeffs(H, out_h)
        tmp = np.zeros((out_h, cur.shape[1], C), dtype=np.uint8)
        for yy in range(out_h):
            y0, n = bounds[yy]
            acc = np.full((cur.shape[1], C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
            for y in range(n):
                acc += cur[y0 + y].astype(np.int64) * int(kk[yy, y])
            tmp[yy] = _clip8(acc)
        cur = tmp
    return cur


def to_tensor_normalize(img_u8, mean, std):
    """ToTensor + Normalize: uint8 [H, W, C] -> float32 [C, H, W]"""
    x = img_u8.astype(np.float32) / np.float32(255.0)
    x = np.transpose(x, (2, 0, 1))
    m = np.asarray(mean, dtype=np.float32).reshape(-1, 1, 1)
    s = np.asarray(std, dtype=np.float32).reshape(-1, 1, 1)
    return ((x - m) / s).astype(np.float32)


def transform(img_u8, out_h, out_w, mean, std):
    return to_tensor_normalize(resize_u8(img_u8, out_h, out_w), mean, std)


def one_hot(labels, n_classes):
    out = np.zeros((len(labels), n_classes), dtype=np.int64)
    out[np.arange(len(labels)), np.asarray(labels)] = 1
    return out
