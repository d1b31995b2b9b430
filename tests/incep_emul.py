"""TEST INFRASTRUCTURE ONLY: a torch restatement of the Inception primitives of include/jck_b200.h (jck_conv_gemm,
jck_im2col, jck_pool3, jck_global_avgpool, jck_resize_norm, jck_inception_score) with the SAME argument meaning, so that
(a) the host graph in jck_generation_b200/inception.py (buffer borders, tap shifts, concat offsets) can be checked against
torchvision on a machine without a GPU, and (b) the CUDA kernels can be checked primitive by primitive on the GPU.
Never imported by the package."""
import torch


def _buf_index(geom, ld, B, H, W):
    """flat element offsets [B, H, W] of logical pixel (b, y, x), channel c_off"""
    Hb, Wb, by, bx, c_off = geom
    b = torch.arange(B).view(B, 1, 1)
    y = torch.arange(H).view(1, H, 1)
    x = torch.arange(W).view(1, 1, W)
    return ((b * Hb + y + by) * Wb + x + bx) * ld + c_off


def conv_gemm(act, lda, w, scale, bias, out, ldc, geom):
    M, N, C, ntaps, Hq, Wq, oy0, ox0, Ho, Wo, Hob, Wob, opy, opx, c_off, relu, out_dtype, rows_a = geom[:18]
    shifts = geom[18:18 + ntaps]
    Cp = 32 if C <= 32 else (C + 63) // 64 * 64
    A = act.view(-1)[:rows_a * lda].view(rows_a, lda)[:, :C].float()
    Wm = w.view(N, ntaps * Cp).float()
    acc = torch.zeros(M, N)
    m = torch.arange(M)
    for t, sh in enumerate(shifts):
        idx = m + sh
        ok = (idx >= 0) & (idx < rows_a)
        At = torch.zeros(M, C)
        At[ok] = A[idx[ok]]
        acc += At @ Wm[:, t * Cp:t * Cp + C].t()
    if scale is not None:
        acc = acc * scale.view(1, N)
    if bias is not None:
        acc = acc + bias.view(1, N)
    if relu:
        acc = acc.clamp_min(0)
    b = m // (Hq * Wq)
    r = m % (Hq * Wq)
    oy, ox = r // Wq - oy0, r % Wq - ox0
    ok = (oy >= 0) & (oy < Ho) & (ox >= 0) & (ox < Wo)
    orow = ((b * Hob + oy + opy) * Wob + ox + opx) * ldc + c_off
    dst = orow[ok].view(-1, 1) + torch.arange(N).view(1, N)
    hi = acc[ok].reshape(-1).to(out.dtype)
    out.view(-1)[dst.reshape(-1)] = hi
    if len(geom) > 18 + ntaps:                  # split-precision output: low plane geom[18 + ntaps] rows behind
        lo_off = geom[18 + ntaps] * ldc
        out.view(-1)[dst.reshape(-1) + lo_off] = (acc[ok].reshape(-1) - hi.float()).to(out.dtype)


def im2col(x, in_geom, ldx, patches, B, H, W, C, kh, kw, sy, sx, py, px, Ho, Wo, Kp):
    src = x.view(-1)
    base = _buf_index(in_geom, ldx, B, H, W)                     # offsets of (b, y, x)
    P = torch.zeros(B, Ho, Wo, Kp, dtype=x.dtype)
    oy = torch.arange(Ho).view(Ho, 1)
    ox = torch.arange(Wo).view(1, Wo)
    for ky in range(kh):
        for kx in range(kw):
            iy, ix = oy * sy + ky - py, ox * sx + kx - px
            ok = ((iy >= 0) & (iy < H) & (ix >= 0) & (ix < W))
            iyc, ixc = iy.clamp(0, H - 1).expand(Ho, Wo), ix.clamp(0, W - 1).expand(Ho, Wo)
            offs = base[:, iyc, ixc]                               # [B, Ho, Wo]
            vals = src[(offs.unsqueeze(-1) + torch.arange(C)).reshape(-1)].view(B, Ho, Wo, C)
            vals = vals * ok.view(1, Ho, Wo, 1).to(vals.dtype)
            k0 = (ky * kw + kx) * C
            P[..., k0:k0 + C] = vals
    patches.view(-1)[:B * Ho * Wo * Kp] = P.view(-1)


def pool3(x, in_geom, ldx, out, out_geom, ldo, B, H, W, C, stride, pad, Ho, Wo, mode, x_lo=0, out_lo=0):
    src = x.view(-1)
    base = _buf_index(in_geom, ldx, B, H, W)
    idx = (base.unsqueeze(-1) + torch.arange(C)).reshape(-1)
    X = src[idx].float()
    if x_lo:
        X = X + src[idx + x_lo].float()
    X = X.view(B, H, W, C).permute(0, 3, 1, 2)
    if mode == 0:
        Y = torch.nn.functional.max_pool2d(X, 3, stride)
    else:
        Y = torch.nn.functional.avg_pool2d(X, 3, stride, pad)
    assert Y.shape[2:] == (Ho, Wo)
    obase = _buf_index(out_geom, ldo, B, Ho, Wo)
    oidx = (obase.unsqueeze(-1) + torch.arange(C)).reshape(-1)
    yv = Y.permute(0, 2, 3, 1).reshape(-1)
    hi = yv.to(out.dtype)
    out.view(-1)[oidx] = hi
    if out_lo:
        out.view(-1)[oidx + out_lo] = (yv - hi.float()).to(out.dtype)


def global_avgpool(x, out_f32, out_bf16, B, HW, C, x_lo=0, out_lo=0):
    v = x.view(-1)[:B * HW * C].float()
    if x_lo:
        v = v + x.view(-1)[x_lo:x_lo + B * HW * C].float()
    m = v.view(B, HW, C).mean(1)
    if out_f32 is not None:
        out_f32.view(-1)[:B * C] = m.reshape(-1)
    if out_bf16 is not None:
        hi = m.reshape(-1).to(out_bf16.dtype)
        out_bf16.view(-1)[:B * C] = hi
        if out_lo:
            out_bf16.view(-1)[out_lo:out_lo + B * C] = (m.reshape(-1) - hi.float()).to(out_bf16.dtype)


def resize_norm(x_nchw, out_nhwc, B, C, Hi, Wi, Ho, Wo, ldo, a, b, mean3, std3):
    y = torch.nn.functional.interpolate(x_nchw.float(), size=(Ho, Wo), mode="bilinear", align_corners=False)
    mean = torch.tensor(mean3[:C]).view(1, C, 1, 1)
    std = torch.tensor(std3[:C]).view(1, C, 1, 1)
    y = (a * y + b - mean) / std
    o = torch.zeros(B, Ho, Wo, ldo)
    o[..., :C] = y.permute(0, 2, 3, 1)
    out_nhwc.view(-1)[:o.numel()] = o.view(-1).to(out_nhwc.dtype)


def stem_patches(x_nchw, patches, B, Hi, Wi, Hr, Wr, a, b, mean3, std3, patches_lo=0):
    Ho, Wo = (Hr - 3) // 2 + 1, (Wr - 3) // 2 + 1
    if patches_lo:                              # split precision: the resized image in fp32, its hi / lo bf16 planes
        f32 = torch.zeros(B * Hr * Wr * 4, dtype=torch.float32)
        resize_norm(x_nchw, f32, B, 3, Hi, Wi, Hr, Wr, 4, a, b, mean3, std3)
        hi = f32.to(patches.dtype)
        lo = (f32 - hi.float()).to(patches.dtype)
        im2col(hi, [Hr, Wr, 0, 0, 0], 4, patches, B, Hr, Wr, 3, 3, 3, 2, 2, 0, 0, Ho, Wo, 32)
        im2col(lo, [Hr, Wr, 0, 0, 0], 4, patches.view(-1)[patches_lo:], B, Hr, Wr, 3, 3, 3, 2, 2, 0, 0, Ho, Wo, 32)
        return
    img = torch.zeros(B * Hr * Wr * 4, dtype=patches.dtype)
    resize_norm(x_nchw, img, B, 3, Hi, Wi, Hr, Wr, 4, a, b, mean3, std3)
    im2col(img, [Hr, Wr, 0, 0, 0], 4, patches, B, Hr, Wr, 3, 3, 3, 2, 2, 0, 0, Ho, Wo, 32)


def inception_score(logits, splits, scores):
    n, d = logits.shape
    per = n // splits
    p = torch.softmax(logits.double(), 1)
    for k in range(splits):
        part = p[k * per:(k + 1) * per]
        py = part.mean(0, keepdim=True)
        kl = (part * (part / py).log()).sum(1).mean()
        scores[k] = float(kl.exp())
