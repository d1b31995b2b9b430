"""Per-kernel parity checks through the C ABI against torch CPU operators (the arithmetic the
reference itself runs).  Used by tests/test_gpu_kernels.py, and runnable on its own with every case
in a fresh subprocess under a timeout so that one trapped kernel cannot take the rest down:

    python -m tests.kernel_checks --isolate
"""
import json
import subprocess
import sys

import torch
import torch.nn.functional as F

# (Ca, Cb, Hs): the three GEMM-shaped layer geometries + the image-edge geometry
SHAPES = {"c2": (128, 64, 16), "c3": (256, 128, 8), "c4": (512, 256, 4), "edge": (64, 3, 32),
          # extra geometries of the windowed 64-channel up-conv kernel (conv_up_win_kernel): one channel chunk; a 32 x 32 map
          "w64": (64, 64, 16), "w32": (128, 64, 32)}


def _rel(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return float((got - want).norm() / max(float(want.norm()), 1e-30))


def _mk(shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(*shape, generator=g) * scale
    return t.to(dtype).float() if dtype == torch.bfloat16 else t


def _packed(w4, dtype):
    from jck_generation_b200 import ops
    Ca, Cb = w4.shape[:2]
    wd = torch.empty(Ca * 16 * Cb, dtype=dtype, device="cuda")
    wu = torch.empty(Ca * 16 * Cb, dtype=dtype, device="cuda")
    ops.pack_weights(w4.cuda().contiguous(), wd, wu)
    return wd, wu


def check_down(shape, dtype, algo, B=8, groups=1):
    from jck_generation_b200 import ops
    Ca, Cb, Hs = SHAPES[shape]
    x = _mk((B, Cb, 2 * Hs, 2 * Hs), dtype, 1)
    w4 = _mk((Ca, Cb, 4, 4), dtype, 2, 0.05)
    want = F.conv2d(x, w4, stride=2, padding=1)
    wd, _ = _packed(w4, dtype)
    xin = x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    out = torch.full((B, Hs, Hs, Ca), float("nan"), dtype=dtype, device="cuda")
    stats = torch.zeros(groups, 2 * Ca, device="cuda")
    ops.conv_down(xin, wd, out, stats, Ca, Cb, ipg=B // groups, algo=algo)
    torch.cuda.synchronize()
    per = B // groups
    ws = torch.stack([torch.cat([want[g * per:(g + 1) * per].sum((0, 2, 3)),
                                 (want[g * per:(g + 1) * per] ** 2).sum((0, 2, 3))]) for g in range(groups)])
    return {"out": _rel(out.float().permute(0, 3, 1, 2), want), "stats": _rel(stats, ws)}


def check_up(shape, dtype, algo, B=8, groups=1, stats=True):
    from jck_generation_b200 import ops
    Ca, Cb, Hs = SHAPES[shape]
    x = _mk((B, Ca, Hs, Hs), dtype, 3)
    w4 = _mk((Ca, Cb, 4, 4), dtype, 4, 0.05)
    want = F.conv_transpose2d(x, w4, stride=2, padding=1)
    _, wu = _packed(w4, dtype)
    xin = x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    out = torch.full((B, 2 * Hs, 2 * Hs, Cb), float("nan"), dtype=dtype, device="cuda")
    if not stats:                # the input-gradient use: no BatchNorm statistics
        ops.conv_up(xin, wu, out, None, Ca, Cb, ipg=B, algo=algo)
        torch.cuda.synchronize()
        return {"out": _rel(out.float().permute(0, 3, 1, 2), want)}
    stats = torch.zeros(groups, 2 * Cb, device="cuda")
    ops.conv_up(xin, wu, out, stats, Ca, Cb, ipg=B // groups, algo=algo)
    torch.cuda.synchronize()
    per = B // groups
    ws = torch.stack([torch.cat([want[g * per:(g + 1) * per].sum((0, 2, 3)),
                                 (want[g * per:(g + 1) * per] ** 2).sum((0, 2, 3))]) for g in range(groups)])
    return {"out": _rel(out.float().permute(0, 3, 1, 2), want), "stats": _rel(stats, ws)}


def _bnbwd_want(da, y, C, groups, slope):
    """torch fp32 restatement of the fused epilogue: g = da * act'(pre), sums = (sum g, sum g * xhat) per group.
    da, y: NCHW fp32.  Returns (g, sums[groups][2C], scale_shift, mean_rstd) with gamma / beta drawn here."""
    per = da.shape[0] // groups
    gamma = 1.0 + 0.1 * _mk((C,), torch.float32, 31)
    beta = 0.1 * _mk((C,), torch.float32, 32)
    gs, sums, ss, mr = [], [], [], []
    for g in range(groups):
        yy, dd = y[g * per:(g + 1) * per], da[g * per:(g + 1) * per]
        mean, var = yy.mean((0, 2, 3)), yy.var((0, 2, 3), unbiased=False)
        rstd = (var + 1e-5).rsqrt()
        sc, sh = gamma * rstd, beta - mean * gamma * rstd
        pre = yy * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)
        gg = torch.where(pre > 0, dd, dd * slope)
        xhat = (yy - mean.view(1, -1, 1, 1)) * rstd.view(1, -1, 1, 1)
        gs.append(gg)
        sums.append(torch.cat([gg.sum((0, 2, 3)), (gg * xhat).sum((0, 2, 3))]))
        ss.append(torch.cat([sc, sh]))
        mr.append(torch.cat([mean, rstd]))
    return torch.cat(gs), torch.stack(sums), torch.stack(ss), torch.stack(mr)


def check_bnbwd(kind, shape, B=8, groups=1, slope=0.2):
    """jck_conv_up_bnbwd / jck_conv_down_bnbwd vs conv + torch BatchNorm-backward reduction."""
    from jck_generation_b200 import ops
    dtype = torch.bfloat16
    if kind == "up":
        Ca, Cb, Hs = SHAPES[shape]
        x = _mk((B, Ca, Hs, Hs), dtype, 3)
        w4 = _mk((Ca, Cb, 4, 4), dtype, 4, 0.05)
        da = F.conv_transpose2d(x, w4, stride=2, padding=1)
        C, Ho = Cb, 2 * Hs
    elif kind == "down":
        Ca, Cb, Hs = SHAPES[shape]
        x = _mk((B, Cb, 2 * Hs, 2 * Hs), dtype, 1)
        w4 = _mk((Ca, Cb, 4, 4), dtype, 2, 0.05)
        da = F.conv2d(x, w4, stride=2, padding=1)
        C, Ho = Ca, Hs
    else:
        raise ValueError(kind)
    y = _mk((B, C, Ho, Ho), dtype, 7)
    g_want, s_want, ss, mr = _bnbwd_want(da, y, C, groups, slope)
    out = torch.full((B, Ho, Ho, C), float("nan"), dtype=dtype, device="cuda")
    sums = torch.zeros(groups, 2 * C, device="cuda")
    y_dev = y.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    ss, mr = ss.cuda().contiguous(), mr.cuda().contiguous()
    wd, wu = _packed(w4, dtype)
    xin = x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    if kind == "up":
        ops.conv_up_bnbwd(xin, wu, y_dev, ss, mr, slope, out, sums, Ca, Cb, ipg=B // groups)
    else:
        ops.conv_down_bnbwd(xin, wd, y_dev, ss, mr, slope, out, sums, Ca, Cb, ipg=B // groups)
    torch.cuda.synchronize()
    return {"out": _rel(out.float().permute(0, 3, 1, 2), g_want), "stats": _rel(sums, s_want)}


def check_wgrad(shape, dtype, algo, B=8):
    from jck_generation_b200 import ops
    Ca, Cb, Hs = SHAPES[shape]
    small = _mk((B, Ca, Hs, Hs), dtype, 5)
    large = _mk((B, Cb, 2 * Hs, 2 * Hs), dtype, 6)
    want = torch.nn.grad.conv2d_weight(large, (Ca, Cb, 4, 4), small, stride=2, padding=1)
    s = small.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    l = large.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    nbytes = ops.wgrad_workspace_bytes(B, Hs, Hs, Ca, Cb, dtype, algo)
    ws = torch.empty(nbytes // 4, device="cuda")
    dw = torch.full((Ca, Cb, 4, 4), float("nan"), device="cuda")
    ops.conv_wgrad(s, l, dw, ws, Ca, Cb, accumulate=False, algo=algo)
    ops.conv_wgrad(s, l, dw, ws, Ca, Cb, accumulate=True, algo=algo)
    torch.cuda.synchronize()
    return {"dw": _rel(dw, 2 * want)}


def _to_p4(x_nchw):
    """NCHW fp32 -> zero-bordered 4-channel bf16 image layout [B][H+2][W+2][4] on the GPU."""
    B, C, H, W = x_nchw.shape
    out = torch.zeros(B, H + 2, W + 2, 4, dtype=torch.bfloat16, device="cuda")
    out[:, 1:-1, 1:-1, :C] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16).cuda()
    return out


def check_edge(which, B=8, nc=3, groups=None):
    """tcgen05 image-edge kernels (Ca = 64, nc image channels, Hs = 32) against torch CPU convolutions."""
    from jck_generation_b200 import ops
    Ca, Hs = 64, 32
    dt = torch.bfloat16
    w4 = _mk((Ca, nc, 4, 4), dt, 31, 0.05)
    wde = torch.empty(Ca * 64, dtype=dt, device="cuda"); wu9 = torch.empty(16 * 9 * Ca, dtype=dt, device="cuda")
    ops.pack_weights_edge(w4.cuda().contiguous(), wde, wu9)
    if which == "down":
        x = _mk((B, nc, 64, 64), dt, 32)
        want = F.conv2d(x, w4, stride=2, padding=1)
        out = torch.full((B, Hs, Hs, Ca), float("nan"), dtype=dt, device="cuda")
        if groups is None:
            groups = 2 if B % 2 == 0 else 1     # two BatchNorm groups when the batch splits evenly
        stats = torch.zeros(groups, 2 * Ca, device="cuda")
        ops.edge_down_img(_to_p4(x), wde, out, stats, Ca, ipg=B // groups)
        torch.cuda.synchronize()
        per = B // groups
        ws = torch.stack([torch.cat([want[g * per:(g + 1) * per].sum((0, 2, 3)), (want[g * per:(g + 1) * per] ** 2).sum((0, 2, 3))])
                          for g in range(groups)])
        return {"out": _rel(out.float().permute(0, 3, 1, 2), want), "stats": _rel(stats, ws)}
    if which in ("up", "upscatter"):
        x = _mk((B, Ca, Hs, Hs), dt, 33)
        want = F.conv_transpose2d(x, w4, stride=2, padding=1)
        img = torch.zeros(B, 66, 66, 4, dtype=dt, device="cuda")
        if which == "up":
            ops.edge_up(x.permute(0, 2, 3, 1).contiguous().to(dt).cuda(), wu9, img, Ca)
        else:
            ops.edge_up_scatter(x.permute(0, 2, 3, 1).contiguous().to(dt).cuda(), wde, img, Ca)
        torch.cuda.synchronize()
        border = float(img[:, 0].abs().sum() + img[:, -1].abs().sum() + img[:, :, 0].abs().sum() + img[:, :, -1].abs().sum()
                       + img[..., nc:].abs().sum())
        return {"out": _rel(img[:, 1:-1, 1:-1, :nc].float().permute(0, 3, 1, 2), want), "border": border}
    small = _mk((B, Ca, Hs, Hs), dt, 34)
    large = _mk((B, nc, 64, 64), dt, 35)
    want = torch.nn.grad.conv2d_weight(large, (Ca, nc, 4, 4), small, stride=2, padding=1)
    ws = torch.empty(ops.edge_wgrad_workspace_bytes(B, Hs, Hs, Ca) // 4, device="cuda")
    dw = torch.full((Ca, nc, 4, 4), float("nan"), device="cuda")
    s = small.permute(0, 2, 3, 1).contiguous().to(dt).cuda()
    img = _to_p4(large)
    ops.edge_wgrad_img(s, img, dw, ws, Ca, nc, False)
    ops.edge_wgrad_img(s, img, dw, ws, Ca, nc, True)
    torch.cuda.synchronize()
    return {"dw": _rel(dw, 2 * want)}


def check_bn(dtype, C=128, B=8, H=16, groups=2):
    """stats -> finalize -> apply -> backward against F.batch_norm + leaky_relu autograd, per group."""
    from jck_generation_b200 import ops
    y = _mk((B, C, H, H), torch.float32, 7) * 1.5 + 0.3
    da = _mk((B, C, H, H), dtype, 8)
    if dtype == torch.bfloat16:
        y = y.to(dtype).float()
    gamma = 1 + 0.1 * _mk((C,), torch.float32, 9)
    beta = 0.1 * _mk((C,), torch.float32, 10)
    per = B // groups
    rm, rv = torch.zeros(C), torch.ones(C)
    want_a, want_dy, want_dg, want_db = [], [], torch.zeros(C), torch.zeros(C)
    for g in range(groups):
        yy = y[g * per:(g + 1) * per].clone().requires_grad_(True)
        gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        a = F.leaky_relu(F.batch_norm(yy, rm, rv, gm, bt, True, 0.1, 1e-5), 0.2)
        a.backward(da[g * per:(g + 1) * per])
        want_a.append(a.detach()); want_dy.append(yy.grad)
        want_dg += gm.grad; want_db += bt.grad
    want_a, want_dy = torch.cat(want_a), torch.cat(want_dy)

    yn = y.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    dn = da.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    yf = yn.float()
    stats = torch.stack([torch.cat([yf[g * per:(g + 1) * per].sum((0, 1, 2)), (yf[g * per:(g + 1) * per] ** 2).sum((0, 1, 2))])
                         for g in range(groups)]).contiguous()
    ss = torch.empty(groups, 2 * C, device="cuda"); mr = torch.empty(groups, 2 * C, device="cuda")
    rmd, rvd = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    gd, bd = gamma.cuda(), beta.cuda()
    ops.bn_finalize(stats, gd, bd, rmd, rvd, nbt, ss, mr, C, groups, per * H * H)
    # every traversal order of the streaming passes (coherent front ascending / descending on tensors >= 24 MB, slabs):
    # the worst error over the three is reported, and the three must agree with each other to summation-order noise
    out, first = {}, None
    for order in (ops.ORDER_DESC, ops.ORDER_ASC, ops.ORDER_SLAB):
        a = torch.empty_like(yn)
        ops.bn_act_fwd(yn, ss, a, C, groups, 0.2, order=order)
        sums = torch.zeros(groups, 2 * C, device="cuda")
        ops.bn_act_bwd_reduce(dn, yn, ss, mr, sums, C, groups, 0.2, order=order)
        dy = torch.empty_like(yn)
        ops.bn_act_bwd_apply(dn, yn, ss, mr, gd, sums, dy, C, groups, per * H * H, 0.2, order=order)
        dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        ops.bn_param_grad(sums, dg, db, C, groups, False)
        torch.cuda.synchronize()
        e = {"a": _rel(a.float().permute(0, 3, 1, 2), want_a), "dy": _rel(dy.float().permute(0, 3, 1, 2), want_dy),
             "dgamma": _rel(dg, want_dg), "dbeta": _rel(db, want_db)}
        if first is None:
            first = (a, sums)
        else:
            e["exact_order_a"] = int((a != first[0]).sum())                            # same arithmetic: bit-identical
            e["order_sums"] = _rel(sums, first[1])                                     # atomics in another order
        for k, v in e.items():
            out[k] = max(out.get(k, 0.0), v)
    out.update({"running_mean": _rel(rmd, rm), "running_var": _rel(rvd, rv), "nbt": abs(int(nbt) - groups)})
    return out


def check_head(dtype, B=8, C4=512):
    from jck_generation_b200 import ops
    K = 16 * C4
    a4 = _mk((B, C4, 4, 4), dtype, 11, 0.5)
    w = _mk((1, C4, 4, 4), dtype, 12, 0.02)
    a4r = a4.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    p = torch.sigmoid(F.conv2d(a4r, wr)).view(-1)
    loss = F.binary_cross_entropy(p, torch.full((B,), 0.9))
    loss.backward()
    an = a4.permute(0, 2, 3, 1).contiguous().to(dtype).cuda().view(B, K)
    w5 = torch.empty(K, dtype=dtype, device="cuda")
    ops.pack_head(w.cuda().contiguous(), w5)
    prob = torch.empty(B, device="cuda"); scal = torch.zeros(2, device="cuda")
    ops.head_fwd(an, w5, prob, 0.9, scal)
    da4 = torch.empty(B, K, dtype=dtype, device="cuda"); dw5 = torch.zeros(K, device="cuda")
    ops.head_bwd(prob, 0.9, w5, an, da4, dw5, 0, False)
    dw4 = torch.zeros(1, C4, 4, 4, device="cuda")
    ops.unpack_head_grad(dw5, dw4, False)
    torch.cuda.synchronize()
    return {"prob": _rel(prob, p.detach()), "loss": abs(float(scal[0]) - float(loss)) / abs(float(loss)),
            "da4": _rel(da4.float().view(B, 4, 4, C4).permute(0, 3, 1, 2), a4r.grad), "dw5": _rel(dw4, wr.grad)}


def check_fc(dtype, B=8, K=100, C=512):
    from jck_generation_b200 import ops
    z = _mk((B, K, 1, 1), torch.float32, 13)
    w4 = _mk((K, C, 4, 4), dtype, 14, 0.05)
    want = F.conv_transpose2d(z, w4, stride=1, padding=0)               # [B,C,4,4]
    dy = _mk((B, C, 4, 4), dtype, 15)
    want_dw = torch.einsum("bk,bcyx->kcyx", z.view(B, K), dy)
    wfc = torch.empty(16 * C, K, dtype=dtype, device="cuda")
    ops.pack_fc(w4.cuda().contiguous(), wfc)
    out = torch.empty(B, 4, 4, C, dtype=dtype, device="cuda"); stats = torch.zeros(1, 2 * C, device="cuda")
    ops.fc_fwd(z.view(B, K).cuda().contiguous(), wfc, out, stats, C)
    dwfc = torch.empty(16 * C, K, device="cuda")
    ops.fc_wgrad(dy.permute(0, 2, 3, 1).contiguous().to(dtype).cuda().view(B, 16 * C), z.view(B, K).cuda().contiguous(), dwfc)
    dw4 = torch.zeros(K, C, 4, 4, device="cuda")
    ops.unpack_fc_grad(dwfc, dw4, False)
    torch.cuda.synchronize()
    ws = torch.cat([want.sum((0, 2, 3)), (want ** 2).sum((0, 2, 3))])
    return {"out": _rel(out.float().permute(0, 3, 1, 2), want), "stats": _rel(stats.view(-1), ws), "dw": _rel(dw4, want_dw)}


def check_fc_tc(B=512, K=100, C=512):
    """G.conv1 as the bf16 engine runs it (engine.GeneratorEngine.forward / backward): z cast to bf16 rows, the weight as the
    MN-major operand of jck_gemm_tc, BatchNorm statistics in the epilogue; weight gradient = split-K GEMM over the batch."""
    from jck_generation_b200 import ops
    bf = torch.bfloat16
    z = _mk((B, K, 1, 1), bf, 13)
    w4 = _mk((K, C, 4, 4), bf, 14, 0.05)
    want = F.conv_transpose2d(z, w4, stride=1, padding=0)               # [B,C,4,4]
    dy = _mk((B, C, 4, 4), bf, 15)
    want_dw = torch.einsum("bk,bcyx->kcyx", z.view(B, K), dy)
    N1, Kp = 16 * C, (K + 7) // 8 * 8
    w_fc = torch.empty(K, N1, dtype=bf, device="cuda")
    ops.pack_fc_t(w4.cuda().contiguous(), w_fc)
    zb = torch.empty(B, Kp, dtype=bf, device="cuda")
    ops.cast_rows_bf16(z.view(B, K).cuda().contiguous(), zb)
    y1 = torch.full((B, 4, 4, C), float("nan"), dtype=bf, device="cuda")
    stats = torch.zeros(1, 2 * C, device="cuda")
    ops.gemm_tc(zb, 0, Kp, w_fc, 1, N1, y1.view(B, N1), B, N1, K, stats=stats, stats_channels=C)
    dyn = dy.permute(0, 2, 3, 1).contiguous().to(bf).cuda().view(B, N1)
    dw_fc = torch.empty(K, N1, device="cuda")
    nbytes = ops.gemm_tc_workspace_bytes(K, N1, B)
    ws = torch.empty(max(nbytes, 4) // 4, device="cuda")
    ops.gemm_tc(zb, 1, Kp, dyn, 1, N1, dw_fc, K, N1, B, workspace=ws)
    dw4 = torch.zeros(K, C, 4, 4, device="cuda")
    ops.unpack_fc_grad_t(dw_fc, dw4, False)
    torch.cuda.synchronize()
    wst = torch.cat([want.sum((0, 2, 3)), (want ** 2).sum((0, 2, 3))])
    return {"out": _rel(y1.float().permute(0, 3, 1, 2), want), "stats": _rel(stats.view(-1), wst), "dw": _rel(dw4, want_dw)}


def check_misc():
    """image edge, generator edge, gp, adam, rng."""
    from jck_generation_b200 import ops
    B, C = 4, 3
    x, m, x2 = _mk((B, C, 64, 64), torch.float32, 16), _mk((B, C, 64, 64), torch.float32, 17), _mk((B, C, 64, 64), torch.float32, 18)
    al = torch.rand(B, generator=torch.Generator().manual_seed(19))
    want = al.view(B, 1, 1, 1) * (0.9 * x + 0.1 * m) + (1 - al.view(B, 1, 1, 1)) * x2
    o1 = torch.empty(B, 64, 64, C, device="cuda"); o2 = torch.empty(B, C, 64, 64, device="cuda")
    ops.prep_image(x.cuda(), out_nhwc=o1, m1=m.cuda(), a1=0.9, b1=0.1, x2=x2.cuda(), alpha=al.cuda(), out_nchw=o2)
    res = {"prep_nhwc": _rel(o1.permute(0, 3, 1, 2), want), "prep_nchw": _rel(o2, want)}
    back = torch.empty(B, C, 64, 64, device="cuda")
    ops.nhwc_to_nchw(o1, back)
    res["nhwc_to_nchw"] = _rel(back, want)
    y5 = _mk((B, 64, 64, C), torch.float32, 20).cuda()
    fr, fm = torch.empty(B, C, 64, 64, device="cuda"), torch.empty(B, C, 64, 64, device="cuda")
    mn = torch.empty(B, 64, 64, C, device="cuda")
    ops.g_out_fwd(y5, m.cuda(), 0.9, 0.1, fr, fm, mn, (B, C, 64, 64))
    t = torch.tanh(y5.cpu().permute(0, 3, 1, 2))
    res["g_out_raw"] = _rel(fr, t); res["g_out_mix"] = _rel(fm, 0.9 * t + 0.1 * m); res["g_out_mix_nhwc"] = _rel(mn.permute(0, 3, 1, 2), 0.9 * t + 0.1 * m)
    dm = _mk((B, 64, 64, C), torch.float32, 21).cuda(); dy5 = torch.empty_like(dm)
    ops.g_out_bwd(dm, fr, 0.9, dy5)
    res["g_out_bwd"] = _rel(dy5.permute(0, 3, 1, 2), 0.9 * dm.cpu().permute(0, 3, 1, 2) * (1 - t * t))
    sc = torch.zeros(2, device="cuda")
    ops.gp_penalty(dm, sc)
    wantgp = ((dm.cpu().reshape(B, -1).norm(2, dim=1) - 1) ** 2).mean()
    res["gp"] = abs(float(sc[0]) - float(wantgp)) / float(wantgp)
    # JCK_IMG_P4 (bf16) quad path, and the noise drawn in registers == randn + the memory path, bit for bit
    bf = torch.bfloat16
    xb, ctr = x.cuda(), torch.full((1,), 77, dtype=torch.int64, device="cuda")
    noise = torch.empty(B, C, 64, 64, device="cuda")
    ops.randn(noise, 99, 5, ctr)
    pa, pb = torch.zeros(B, 66, 66, 4, dtype=bf, device="cuda"), torch.zeros(B, 66, 66, 4, dtype=bf, device="cuda")
    na, nb = torch.empty(B, C, 64, 64, device="cuda"), torch.empty(B, C, 64, 64, device="cuda")
    ops.prep_image(xb, out_nhwc=pa, m1=noise, a1=0.9, b1=0.1, out_nchw=na, layout=ops.IMG_P4)
    ops.prep_image_rng(xb, 99, 5, ctr, 0.9, 0.1, out_nhwc=pb, out_nchw=nb, layout=ops.IMG_P4)
    res["exact_prep_rng"] = float(bool((pa != pb).any()) or bool((na != nb).any()))
    wantp = 0.9 * x + 0.1 * noise.cpu()
    res["prep_p4_bf16"] = _rel(pa[:, 1:65, 1:65, :3].permute(0, 3, 1, 2).float(), wantp)
    inner = torch.zeros(B, 66, 66, 4, dtype=torch.bool, device="cuda")
    inner[:, 1:65, 1:65, :3] = True
    res["border"] = float(bool((pa[~inner] != 0).any()))
    y5p = torch.zeros(B, 66, 66, 4, dtype=bf, device="cuda")
    y5p[:, 1:65, 1:65, :3] = y5.to(bf)
    outs = []
    for drawn in (False, True):
        fr2, fm2 = torch.empty(B, C, 64, 64, device="cuda"), torch.empty(B, C, 64, 64, device="cuda")
        mp = torch.zeros(B, 66, 66, 4, dtype=bf, device="cuda")
        if drawn:
            ops.g_out_fwd_rng(y5p, 99, 5, ctr, 0.9, 0.1, fr2, fm2, mp, (B, C, 64, 64), layout=ops.IMG_P4)
        else:
            ops.g_out_fwd(y5p, noise, 0.9, 0.1, fr2, fm2, mp, (B, C, 64, 64), layout=ops.IMG_P4)
        outs.append((fr2, fm2, mp))
    res["exact_g_out_rng"] = float(any(bool((a != b).any()) for a, b in zip(*outs)))
    tp = torch.tanh(y5.to(bf).float().cpu().permute(0, 3, 1, 2))
    res["g_out_p4_raw"] = _rel(outs[0][0], tp)
    res["g_out_p4_mix_bf16"] = _rel(outs[0][2][:, 1:65, 1:65, :3].permute(0, 3, 1, 2).float(), 0.9 * tp + 0.1 * noise.cpu())
    dmp = torch.zeros(B, 66, 66, 4, dtype=bf, device="cuda")
    dmp[:, 1:65, 1:65, :3] = dm.to(bf)
    dyp = torch.zeros_like(dmp)
    ops.g_out_bwd(dmp, outs[0][0], 0.9, dyp, layout=ops.IMG_P4)
    res["g_out_bwd_p4_bf16"] = _rel(dyp[:, 1:65, 1:65, :3].permute(0, 3, 1, 2).float(),
                                    0.9 * dm.to(bf).float().cpu().permute(0, 3, 1, 2) * (1 - tp * tp))
    res["border2"] = float(bool((dyp[~inner] != 0).any()) or bool((outs[1][2][~inner] != 0).any()))
    # Adam, 3 steps against torch.optim.Adam
    p0, g0 = _mk((1000,), torch.float32, 22), [_mk((1000,), torch.float32, 23 + i) for i in range(3)]
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=2e-4, betas=(0.5, 0.999))
    pd, md, vd = p0.cuda(), torch.zeros(1000, device="cuda"), torch.zeros(1000, device="cuda")
    stepc = torch.zeros(1, dtype=torch.int32, device="cuda")
    for gi in g0:
        pt.grad = gi.clone(); opt.step()
        ops.adam(pd, gi.cuda(), md, vd, 2e-4, 0.5, 0.999, 1e-8, stepc); ops.adam_advance(stepc)
    res["adam"] = _rel(pd - p0.cuda(), pt.detach() - p0)
    # RNG moments
    r = torch.empty(1 << 20, device="cuda"); ctr = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.randn(r, 1234, 1, ctr)
    res["randn_mean"] = abs(float(r.mean())); res["randn_std"] = abs(float(r.std()) - 1)
    ops.rand(r, 1234, 2, ctr)
    res["rand_mean"] = abs(float(r.mean()) - 0.5); res["rand_range"] = float((r.min() < 0) | (r.max() >= 1))
    torch.cuda.synchronize()
    return res


def check_gemm():
    """jck_gemm_tc: all operand major-ness combinations, ragged M / N / K, split-K, accumulate, bf16 output."""
    from jck_generation_b200 import ops
    bf = torch.bfloat16
    res = {}

    def operand(rows, K, mn, seed):
        x = _mk((rows, K), torch.float32, seed).to(bf)            # logical [rows][K]
        pad = lambda v: (v + 7) // 8 * 8
        if mn:
            st = torch.zeros(K, pad(rows), dtype=bf)
            st[:, :rows] = x.t()
        else:
            st = torch.zeros(rows, pad(K), dtype=bf)
            st[:, :K] = x
        return x.float(), st.cuda(), st.shape[1]

    for name, (M, N, K, amn, bmn, cdt, acc) in {
            "fwd_splitk": (300, 256, 8192, 0, 0, torch.float32, False),
            "fwd_acc": (300, 256, 1000, 0, 0, torch.float32, True),
            "dgrad_bf16": (300, 1000, 256, 0, 1, bf, False),
            "wgrad_acc": (256, 1000, 300, 1, 1, torch.float32, True),
            "mn_k": (130, 70, 200, 1, 0, torch.float32, False)}.items():
        a, A, lda = operand(M, K, amn, 40)
        b, Bm, ldb = operand(N, K, bmn, 41)
        c0 = _mk((M, N), torch.float32, 42)
        C = (c0.clone() if acc else torch.full((M, N), 7.0)).to(cdt).cuda()
        nbytes = ops.gemm_tc_workspace_bytes(M, N, K)
        ws = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device="cuda")
        ops.gemm_tc(A, amn, lda, Bm, bmn, ldb, C, M, N, K, accumulate=acc, workspace=ws)
        want = a @ b.t() + (c0 if acc else 0)
        res["gemm_" + name] = _rel(C.float(), want)
    torch.cuda.synchronize()
    return res


def _alg(name):
    from jck_generation_b200 import ops
    return {"simt": ops.ALGO_SIMT, "tc": ops.ALGO_TC, "auto": ops.ALGO_AUTO}[name]


def _dt(name):
    return torch.float32 if name == "f32" else torch.bfloat16


def all_cases():
    cases = []
    for op in ("down", "up", "wgrad"):
        for shape in ("c2", "c3", "c4", "edge"):
            cases.append((op, shape, "f32", "simt", 8))
            cases.append((op, shape, "bf16", "simt", 8))
            if shape != "edge":
                cases.append((op, shape, "bf16", "tc", 8))
                cases.append((op, shape, "bf16", "tc", 3))      # ragged: batch not a multiple of the tile
    cases += [("up", "w64", "bf16", "tc", 5), ("up", "w32", "bf16", "tc", 3), ("up_groups", "c2", "bf16", "tc", 6),
              ("up_groups", "w32", "bf16", "tc", 4), ("up_nostats", "c2", "bf16", "tc", 9)]
    cases += [("down_groups", "c3", "bf16", "tc", 8), ("up_groups", "c4", "bf16", "tc", 16),
              ("down_groups", "c4", "f32", "simt", 6)]
    cases += [("edge_down", "-", "bf16", "tc", 8), ("edge_down", "-", "bf16", "tc", 3), ("edge_down", "-", "bf16", "tc", 150),
              ("edge_up", "-", "bf16", "tc", 8), ("edge_up", "-", "bf16", "tc", 5), ("edge_wgrad", "-", "bf16", "tc", 8),
              ("edge_wgrad", "-", "bf16", "tc", 3), ("edge_wgrad", "-", "bf16", "tc", 150),
              ("edge_upscatter", "-", "bf16", "tc", 8), ("edge_upscatter", "-", "bf16", "tc", 5), ("edge_upscatter", "-", "bf16", "tc", 150),
              ("edge_upscatter1", "-", "bf16", "tc", 4),
              ("edge_down1", "-", "bf16", "tc", 4), ("edge_up1", "-", "bf16", "tc", 4), ("edge_wgrad1", "-", "bf16", "tc", 4)]
    cases += [("bnbwd_up", "c2", "bf16", "tc", 8), ("bnbwd_up", "c3", "bf16", "tc", 8), ("bnbwd_up", "c4", "bf16", "tc", 3),
              ("bnbwd_down", "c2", "bf16", "tc", 8), ("bnbwd_down", "c3", "bf16", "tc", 3), ("bnbwd_down", "c4", "bf16", "tc", 16)]
    cases += big_cases()
    cases += [("bn", "-", "f32", "-", 8), ("bn", "-", "bf16", "-", 8), ("head", "-", "f32", "-", 8),
              ("head", "-", "bf16", "-", 8), ("fc", "-", "f32", "-", 8), ("fc", "-", "bf16", "-", 8),
              ("misc", "-", "f32", "-", 4), ("gemm", "-", "bf16", "tc", 0)]
    return cases


def big_cases():
    """The sizes the benchmark runs (BASELINE configs[2]: 512 images per GPU).  At these sizes every tcgen05 kernel walks
    several tiles per CTA (persistent tile loop, TMEM double buffering, ring phase carry-over, register-resident
    BatchNorm partials flushed on a group change, split-K ranges of the weight gradient) -- paths a batch of 8 never
    enters (<= 32 tile pairs < 148 clusters).  1536 = the discriminator's [real | fake | x_hat] forward with three
    BatchNorm groups; 1024 = the backward of its A+B slice (two groups); 512 = the generator and the penalty sweep."""
    cases = []
    for shape in ("c2", "c3", "c4"):
        cases += [("down", shape, "bf16", "tc", 512), ("down_groups3", shape, "bf16", "tc", 1536),
                  ("up", shape, "bf16", "tc", 512), ("up", shape, "bf16", "tc", 1024),
                  ("wgrad", shape, "bf16", "tc", 1024), ("wgrad", shape, "bf16", "tc", 512)]
    cases += [("bnbwd_up", "c2", "bf16", "tc", 1024), ("bnbwd_up", "c2", "bf16", "tc", 512), ("bnbwd_up", "w32", "bf16", "tc", 6),
              ("bnbwd_up", "c3", "bf16", "tc", 1024), ("bnbwd_up", "c4", "bf16", "tc", 1024), ("bnbwd_up", "c4", "bf16", "tc", 512),
              ("bnbwd_down", "c3", "bf16", "tc", 512), ("bnbwd_down", "c4", "bf16", "tc", 512),
              ("edge_down_g3", "-", "bf16", "tc", 1536), ("edge_down", "-", "bf16", "tc", 512),
              ("edge_upscatter", "-", "bf16", "tc", 512), ("edge_wgrad", "-", "bf16", "tc", 1024),
              ("edge_wgrad", "-", "bf16", "tc", 512),
              ("bn_big", "-", "bf16", "-", 1536), ("head_big", "-", "bf16", "-", 512), ("fc_big", "-", "bf16", "-", 512)]
    return cases


def run_case(op, shape, dtype, algo, B):
    if op == "down":
        return check_down(shape, _dt(dtype), _alg(algo), B)
    if op == "up":
        return check_up(shape, _dt(dtype), _alg(algo), B)
    if op == "wgrad":
        return check_wgrad(shape, _dt(dtype), _alg(algo), B)
    if op == "down_groups":
        return check_down(shape, _dt(dtype), _alg(algo), B, groups=2)
    if op == "up_groups":
        return check_up(shape, _dt(dtype), _alg(algo), B, groups=2)
    if op == "up_nostats":
        return check_up(shape, _dt(dtype), _alg(algo), B, stats=False)
    if op == "down_groups3":
        return check_down(shape, _dt(dtype), _alg(algo), B, groups=3)
    if op == "edge_down_g3":
        return check_edge("down", B, nc=3, groups=3)
    if op == "bn_big":
        return check_bn(_dt(dtype), C=128, B=B, H=16, groups=3)
    if op == "head_big":
        return check_head(_dt(dtype), B=B)
    if op == "fc_big":
        return check_fc_tc(B)
    if op.startswith("bnbwd_"):
        kind = op[6:]
        groups = 2 if B % 2 == 0 else 1
        return check_bnbwd(kind, shape, B, groups=groups, slope=0.0 if kind != "up" else 0.2)
    if op.startswith("edge_"):
        return check_edge(op[5:].rstrip("1"), B, nc=1 if op.endswith("1") else 3)
    if op == "bn":
        return check_bn(_dt(dtype))
    if op == "head":
        return check_head(_dt(dtype))
    if op == "fc":
        return check_fc(_dt(dtype))
    if op == "misc":
        return check_misc()
    if op == "gemm":
        return check_gemm()
    raise ValueError(op)


def tolerance(op, dtype, key):
    if key in ("nbt", "rand_range", "border", "border2") or key.startswith("exact_"):
        return 0.5
    if key.endswith("_bf16"):
        return 4e-3
    if key.startswith("randn") or key.startswith("rand_"):
        return 5e-3
    if dtype == "f32":
        return 2e-5 if key != "adam" else 5e-5
    # bf16 storage: outputs are rounded to 8 bits of mantissa (2^-9 relative per element)
    return {"dw": 2e-5, "dw5": 2e-5, "stats": 2e-4, "loss": 1e-3, "prob": 1e-3, "dgamma": 1e-2, "dbeta": 1e-2,
            "running_mean": 1e-3, "running_var": 1e-3}.get(key, 4e-3)


def main():
    if "--one" in sys.argv:
        spec = json.loads(sys.argv[sys.argv.index("--one") + 1])
        print("RESULT " + json.dumps(run_case(*spec)))
        return
    bad = 0
    cases = all_cases()
    if "--only-big" in sys.argv:
        cases = big_cases()
    if "--match" in sys.argv:                      # e.g. --match up:c2  (op prefix : shape)
        op, _, shape = sys.argv[sys.argv.index("--match") + 1].partition(":")
        cases = [c for c in cases if c[0].startswith(op) and (not shape or c[1] == shape)]
    if "--only-tc" in sys.argv:
        cases = [c for c in cases if c[3] == "tc"]
    if "--no-tc" in sys.argv:
        cases = [c for c in cases if c[3] != "tc"]
    for spec in cases:
        if "--isolate" in sys.argv:
            try:
                out = subprocess.run([sys.executable, "-m", "tests.kernel_checks", "--one", json.dumps(spec)],
                                     capture_output=True, text=True, timeout=120)
                line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
                res = json.loads(line[0][7:]) if line else {"error": (out.stderr or out.stdout)[-400:]}
            except subprocess.TimeoutExpired:
                res = {"error": "timeout"}
        else:
            res = run_case(*spec)
        status = "ok"
        for k, v in res.items():
            if k == "error" or not (v <= tolerance(spec[0], spec[2], k)):
                status = "FAIL"
        bad += status != "ok"
        print(status, spec, {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in res.items()}, flush=True)
    print("failures:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
