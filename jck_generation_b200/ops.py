"""Tensor-level wrappers over the C ABI (include/jck_b200.h).  PyTorch is plumbing here: it owns
device memory and streams; every arithmetic kernel is ours.  Nothing in this module computes on
the CPU or through torch operators."""
import ctypes as _ct

import torch

from . import _lib
from ._lib import JCK_BF16, JCK_F32, ALGO_AUTO, ALGO_SIMT, ALGO_TC, IMG_NHWC, IMG_P4, check  # noqa: F401

_DT = {torch.float32: JCK_F32, torch.bfloat16: JCK_BF16}


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    return _DT[d]


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "jck ops need contiguous CUDA tensors"
    return t.data_ptr()


def _s():
    return torch.cuda.current_stream().cuda_stream


def L():
    return _lib.load()


# ---- image edge ----------------------------------------------------------------------------------
def img_alloc(B, C, H, W, dtype, device, layout):
    """Activation-side image tensor: dense NHWC, or the zero-bordered 4-channel JCK_IMG_P4 layout."""
    if layout == IMG_P4:
        return torch.zeros(B, H + 2, W + 2, 4, dtype=dtype, device=device)
    return torch.empty(B, H, W, C, dtype=dtype, device=device)


def prep_image(x1, out_nhwc=None, m1=None, a1=1.0, b1=0.0, x2=None, alpha=None, out_nchw=None, layout=IMG_NHWC):
    B, C, H, W = x1.shape
    d = dt(out_nhwc) if out_nhwc is not None else JCK_F32
    check(L().jck_prep_image(_p(x1), _p(m1), a1, b1, _p(x2), _p(alpha), _p(out_nhwc), _p(out_nchw),
                             B, C, H, W, layout, d, _s()), "prep_image")


def prep_image_rng(x1, seed, stream_id, counter, a1, b1, out_nhwc=None, out_nchw=None, layout=IMG_NHWC):
    """prep_image with m1 ~ N(0,1) drawn in registers from our Philox stream (never stored)."""
    B, C, H, W = x1.shape
    d = dt(out_nhwc) if out_nhwc is not None else JCK_F32
    check(L().jck_prep_image_rng(_p(x1), seed, stream_id, _p(counter), a1, b1, _p(out_nhwc), _p(out_nchw),
                                 B, C, H, W, layout, d, _s()), "prep_image_rng")


def nhwc_to_nchw(x_nhwc, out_nchw, layout=IMG_NHWC):
    B, C, H, W = out_nchw.shape
    check(L().jck_nhwc_to_nchw_f32(_p(x_nhwc), _p(out_nchw), B, C, H, W, layout, dt(x_nhwc), _s()), "nhwc_to_nchw")


# ---- weights ---------------------------------------------------------------------------------------
def pack_weights(w4, w_down, w_up):
    Ca, Cb = w4.shape[0], w4.shape[1]
    d = dt(w_down if w_down is not None else w_up)
    check(L().jck_pack_weights(_p(w4), _p(w_down), _p(w_up), Ca, Cb, d, _s()), "pack_weights")


def pack_weights_edge(w4, w_down_e, w_up9):
    check(L().jck_pack_weights_edge(_p(w4), _p(w_down_e), _p(w_up9), w4.shape[0], w4.shape[1], _s()), "pack_weights_edge")




def edge_wgrad_img(small, img_p4, dw4, workspace, Ca, nc, accumulate):
    """edge_wgrad reading the JCK_IMG_P4 image itself (no patch matrix)."""
    B, Hs, Ws = small.shape[0], small.shape[1], small.shape[2]
    check(L().jck_edge_wgrad_img(_p(small), _p(img_p4), _p(dw4), _p(workspace), workspace.numel() * workspace.element_size(),
                                 B, Hs, Ws, Ca, nc, int(accumulate), _s()), "edge_wgrad_img")


def edge_down_img(img_p4, w_down_e, out_small, stats, Ca, ipg=0):
    """edge_down reading the JCK_IMG_P4 image itself (no patch matrix)."""
    B, Hs, Ws = out_small.shape[0], out_small.shape[1], out_small.shape[2]
    check(L().jck_edge_down_img(_p(img_p4), _p(w_down_e), _p(out_small), _p(stats), B, Hs, Ws, Ca, ipg, _s()), "edge_down_img")


def edge_up_scatter(x_small, w_down_e, img_p4, Ca):
    """edge_up in scatter form (one read of the activations; needs 32-pixel rows)."""
    B, Hs, Ws = x_small.shape[0], x_small.shape[1], x_small.shape[2]
    check(L().jck_edge_up_scatter(_p(x_small), _p(w_down_e), _p(img_p4), B, Hs, Ws, Ca, _s()), "edge_up_scatter")


def edge_up(x_small, w_up9, img_p4, Ca):
    B, Hs, Ws = x_small.shape[0], x_small.shape[1], x_small.shape[2]
    check(L().jck_edge_up(_p(x_small), _p(w_up9), _p(img_p4), B, Hs, Ws, Ca, _s()), "edge_up")


def edge_wgrad_workspace_bytes(B, Hs, Ws, Ca):
    return int(L().jck_edge_wgrad_workspace_bytes(B, Hs, Ws, Ca))



def pack_fc(w4, w_fc):
    K, C = w4.shape[0], w4.shape[1]
    check(L().jck_pack_fc(_p(w4), _p(w_fc), K, C, dt(w_fc), _s()), "pack_fc")


def unpack_fc_grad(dw_fc, dw4, accumulate):
    K, C = dw4.shape[0], dw4.shape[1]
    check(L().jck_unpack_fc_grad(_p(dw_fc), _p(dw4), K, C, int(accumulate), _s()), "unpack_fc_grad")


def pack_head(w4, w5):
    check(L().jck_pack_head(_p(w4), _p(w5), w4.shape[1], dt(w5), _s()), "pack_head")


def unpack_head_grad(dw5, dw4, accumulate):
    check(L().jck_unpack_head_grad(_p(dw5), _p(dw4), dw4.shape[1], int(accumulate), _s()), "unpack_head_grad")


# ---- convolutions ------------------------------------------------------------------------------------
def conv_down(x_large, w_down, out_small, stats, Ca, Cb, ipg=0, algo=ALGO_AUTO):
    B, Hs, Ws = out_small.shape[0], out_small.shape[1], out_small.shape[2]
    check(L().jck_conv_down(_p(x_large), _p(w_down), _p(out_small), _p(stats), B, Hs, Ws, Ca, Cb, ipg,
                            dt(x_large), algo, _s()), "conv_down")


def conv_up(x_small, w_up, out_large, stats, Ca, Cb, ipg=0, algo=ALGO_AUTO):
    B, Hs, Ws = x_small.shape[0], x_small.shape[1], x_small.shape[2]
    check(L().jck_conv_up(_p(x_small), _p(w_up), _p(out_large), _p(stats), B, Hs, Ws, Ca, Cb, ipg,
                          dt(x_small), algo, _s()), "conv_up")


def conv_up_bnbwd(x_small, w_up, y_saved, scale_shift, mean_rstd, slope, out_g, sums, Ca, Cb, ipg=0):
    """conv_up whose epilogue also does the BatchNorm-backward reduction of the layer below (bf16 / tcgen05)."""
    B, Hs, Ws = x_small.shape[0], x_small.shape[1], x_small.shape[2]
    check(L().jck_conv_up_bnbwd(_p(x_small), _p(w_up), _p(y_saved), _p(scale_shift), _p(mean_rstd), float(slope), _p(out_g),
                                _p(sums), B, Hs, Ws, Ca, Cb, ipg, dt(x_small), _s()), "conv_up_bnbwd")


def conv_down_bnbwd(x_large, w_down, y_saved, scale_shift, mean_rstd, slope, out_g, sums, Ca, Cb, ipg=0):
    B, Hs, Ws = out_g.shape[0], out_g.shape[1], out_g.shape[2]
    check(L().jck_conv_down_bnbwd(_p(x_large), _p(w_down), _p(y_saved), _p(scale_shift), _p(mean_rstd), float(slope),
                                  _p(out_g), _p(sums), B, Hs, Ws, Ca, Cb, ipg, dt(x_large), _s()), "conv_down_bnbwd")



def wgrad_workspace_bytes(B, Hs, Ws, Ca, Cb, dtype, algo=ALGO_AUTO):
    return int(L().jck_conv_wgrad_workspace_bytes(B, Hs, Ws, Ca, Cb, dt(dtype), algo))


def conv_wgrad(small, large, dw4, workspace, Ca, Cb, accumulate, algo=ALGO_AUTO):
    B, Hs, Ws = small.shape[0], small.shape[1], small.shape[2]
    check(L().jck_conv_wgrad(_p(small), _p(large), _p(dw4), _p(workspace), workspace.numel() * workspace.element_size(),
                             B, Hs, Ws, Ca, Cb, int(accumulate), dt(small), algo, _s()), "conv_wgrad")


def fc_fwd(x, w_fc, out, stats, C):
    M, K = x.shape
    N = w_fc.shape[0]
    check(L().jck_fc_fwd(_p(x), _p(w_fc), _p(out), _p(stats), M, N, K, C, dt(w_fc), _s()), "fc_fwd")


def fc_wgrad(dy, x, dw_fc, accumulate=False):
    M, K = x.shape
    N = dw_fc.shape[0]
    check(L().jck_fc_wgrad(_p(dy), _p(x), _p(dw_fc), M, N, K, int(accumulate), dt(dy), _s()), "fc_wgrad")


# ---- BatchNorm + activation ----------------------------------------------------------------------------
def bn_finalize(stats, gamma, beta, running_mean, running_var, nbt, scale_shift, mean_rstd, C, groups, count,
                eps=1e-5, momentum=0.1):
    check(L().jck_bn_finalize(_p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(nbt),
                              _p(scale_shift), _p(mean_rstd), C, groups, float(count), eps, momentum, _s()),
          "bn_finalize")


# ---- peer-memory communicator (SyncBN exchanges) ---------------------------------------------------------------
COMM_MAX_N = 3072
COMM_HANDLE_BYTES = 64


def comm_create(rank, world):
    """-> (opaque communicator, 64-byte CUDA IPC handle of this rank's mailbox)"""
    import ctypes
    comm = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(COMM_HANDLE_BYTES)
    check(L().jck_comm_create(rank, world, ctypes.byref(comm), handle), "comm_create")
    return comm, handle.raw


def comm_connect(comm, all_handles):
    check(L().jck_comm_connect(comm, all_handles), "comm_connect")


def comm_configure(comm, early_dependents):
    check(L().jck_comm_configure(comm, int(early_dependents)), "comm_configure")


def comm_error(comm):
    """True when an exchange on this peer communicator ever timed out waiting for a peer (synchronises with the device)."""
    import ctypes
    flag = ctypes.c_int(0)
    check(L().jck_comm_error(comm, ctypes.byref(flag)), "comm_error")
    return bool(flag.value)


def comm_destroy(comm):
    check(L().jck_comm_destroy(comm), "comm_destroy")


def comm_allreduce_small(comm, t):
    assert t.dtype == torch.float32 and t.numel() <= COMM_MAX_N
    check(L().jck_comm_allreduce_small(comm, _p(t), t.numel(), _s()), "comm_allreduce_small")


def bn_finalize_sync(comm, stats, gamma, beta, running_mean, running_var, nbt, scale_shift, mean_rstd, C, groups, count,
                     eps=1e-5, momentum=0.1):
    check(L().jck_bn_finalize_sync(comm, _p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(nbt),
                                   _p(scale_shift), _p(mean_rstd), C, groups, float(count), eps, momentum, _s()),
          "bn_finalize_sync")


def bn_bwd_sums_sync(comm, sums, dgamma, dbeta, C, groups, accumulate):
    check(L().jck_bn_bwd_sums_sync(comm, _p(sums), _p(dgamma), _p(dbeta), C, groups, int(accumulate), _s()),
          "bn_bwd_sums_sync")


# traversal order of the streaming BatchNorm passes on tensors larger than the L2 (include/jck_b200.h: JCK_ORDER_*)
ORDER_ASC, ORDER_DESC, ORDER_SLAB = 0, 1, 2


def bn_act_fwd(y, scale_shift, a, C, groups, slope, order=ORDER_DESC):
    """default DESC: the producer (a convolution) wrote y ascending, the consumer (the next convolution) reads a ascending"""
    npix = y.numel() // C
    check(L().jck_bn_act_fwd(_p(y), _p(scale_shift), _p(a), npix, C, npix // groups, slope, dt(y), order, _s()), "bn_act_fwd")


def bn_act_bwd_reduce(da, y, scale_shift, mean_rstd, sums, C, groups, slope, order=ORDER_DESC):
    npix = y.numel() // C
    check(L().jck_bn_act_bwd_reduce(_p(da), _p(y), _p(scale_shift), _p(mean_rstd), _p(sums), npix, C,
                                    npix // groups, slope, dt(y), order, _s()), "bn_act_bwd_reduce")


def bn_act_bwd_apply(da, y, scale_shift, mean_rstd, gamma, sums, dy, C, groups, count, slope, order=ORDER_ASC):
    """default ASC: it follows a descending reduce pass over the same two tensors"""
    npix = y.numel() // C
    check(L().jck_bn_act_bwd_apply(_p(da), _p(y), _p(scale_shift), _p(mean_rstd), _p(gamma), _p(sums), _p(dy),
                                   npix, C, npix // groups, float(count), slope, dt(y), order, _s()), "bn_act_bwd_apply")


def bn_param_grad(sums, dgamma, dbeta, C, groups, accumulate):
    check(L().jck_bn_param_grad(_p(sums), _p(dgamma), _p(dbeta), C, groups, int(accumulate), _s()), "bn_param_grad")


# ---- DCGAN head --------------------------------------------------------------------------------------------
def head_fwd(a4, w5, prob, target, scalars):
    B = prob.shape[0]
    K = w5.numel()
    check(L().jck_head_fwd(_p(a4), _p(w5), _p(prob), float(target), _p(scalars), B, K, dt(a4), _s()), "head_fwd")


def head_bwd(prob, target, w5, a4, da4, dw5, mode, accumulate, dprob=None, mean_count=0):
    """mean_count: the batch the BCE mean runs over (0 = these rows; the global batch under data parallelism)."""
    B = prob.shape[0]
    K = w5.numel()
    check(L().jck_head_bwd(_p(prob), _p(dprob), float(target), _p(w5), _p(a4), _p(da4), _p(dw5), B, int(mean_count), K, mode,
                           int(accumulate), dt(da4), _s()), "head_bwd")


# ---- generator output, GP, Adam, RNG ---------------------------------------------------------------------------
def g_out_fwd(y5, noise, a, b, fake_raw, fake_mix, mix_nhwc, shape, layout=IMG_NHWC):
    """shape = (B, C, H, W) of the image (y5 may be in the padded JCK_IMG_P4 layout)."""
    B, C, H, W = shape
    check(L().jck_g_out_fwd(_p(y5), _p(noise), a, b, _p(fake_raw), _p(fake_mix), _p(mix_nhwc), B, C, H, W, layout,
                            dt(y5), _s()), "g_out_fwd")


def g_out_fwd_rng(y5, seed, stream_id, counter, a, b, fake_raw, fake_mix, mix_nhwc, shape, layout=IMG_NHWC):
    B, C, H, W = shape
    check(L().jck_g_out_fwd_rng(_p(y5), seed, stream_id, _p(counter), a, b, _p(fake_raw), _p(fake_mix), _p(mix_nhwc),
                                B, C, H, W, layout, dt(y5), _s()), "g_out_fwd_rng")


def g_out_bwd(dmix, fake_raw, a, dy5, layout=IMG_NHWC):
    B, C, H, W = fake_raw.shape
    check(L().jck_g_out_bwd(_p(dmix), _p(fake_raw), a, _p(dy5), B, C, H, W, layout, dt(dmix), _s()), "g_out_bwd")


def gp_penalty(dx, scalars):
    B = dx.shape[0]
    check(L().jck_gp_penalty(_p(dx), _p(scalars), B, dx.numel() // B, dt(dx), _s()), "gp_penalty")


def adam(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step_count):
    check(L().jck_adam(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), lr, beta1, beta2, eps,
                       _p(step_count), _s()), "adam")


def adam_advance(step_count):
    check(L().jck_adam_advance(_p(step_count), _s()), "adam_advance")


def randn(out, seed, stream_id, counter):
    check(L().jck_randn(_p(out), out.numel(), seed, stream_id, _p(counter), _s()), "randn")


def rand(out, seed, stream_id, counter):
    check(L().jck_rand(_p(out), out.numel(), seed, stream_id, _p(counter), _s()), "rand")


def rng_advance(counter, by):
    check(L().jck_rng_advance(_p(counter), by, _s()), "rng_advance")


# ---- CGAN head / second-order helpers -----------------------------------------------------------------------
def dense(A, sam, sak, B, sbn, sbk, C, M, N, K, accumulate=False):
    """C[m][n] (+)= sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]; C row-major with leading dimension C.shape[-1]."""
    ldc = C.shape[-1] if C.dim() > 1 else N
    check(L().jck_dense(_p(A), dt(A), sam, sak, _p(B), dt(B), sbn, sbk, _p(C), dt(C), ldc, M, N, K, int(accumulate), _s()),
          "dense")


def f32_to_bf16(x, out):
    check(L().jck_f32_to_bf16(_p(x), _p(out), x.numel(), _s()), "f32_to_bf16")


def center_split_bf16(x, mean, hi, lo):
    rows, d = x.shape
    check(L().jck_center_split_bf16(_p(x), _p(mean), _p(hi), _p(lo), rows, d, hi.shape[1], _s()), "center_split_bf16")


def feature_moments(feats, comm=None):
    """Mean [d] and covariance [d, d] (np.cov(rowvar=False): divisor N - 1) of device-resident fp32 features [N, d] on our
    kernels: column sums by jck_dense, centred hi/lo bf16 split, Gram matrix = three jck_gemm_tc products with both
    operands MN-major and fp32 accumulation.  With `comm` every rank passes its shard of the rows: the sums and the
    Gram matrix are all-reduced (N = total rows), every rank gets the same result.  Returns fp32 device tensors."""
    assert feats.is_cuda and feats.dtype == torch.float32 and feats.dim() == 2
    feats = feats.contiguous()
    n_loc, d = feats.shape
    dev = feats.device
    n = n_loc
    if comm is not None and comm.world_size > 1:
        cnt = torch.full((1,), float(n_loc), dtype=torch.float32, device=dev)
        comm.allreduce_sum_(cnt)
        n = int(round(float(cnt.item())))
    ones = torch.ones(n_loc, dtype=torch.float32, device=dev)
    sums = torch.zeros(d, dtype=torch.float32, device=dev)
    dense(ones, 0, 1, feats, 1, d, sums, 1, d, n_loc)                     # column sums
    if comm is not None:
        comm.allreduce_sum_(sums)
    mean = torch.empty(d, dtype=torch.float32, device=dev)
    rowop(ROW_SCALE_ROWS, sums.view(d, 1), torch.full((d,), 1.0 / n, dtype=torch.float32, device=dev), mean.view(d, 1), d, 1)
    ld = (d + 7) // 8 * 8
    hi = torch.empty(n_loc, ld, dtype=torch.bfloat16, device=dev)
    lo = torch.empty(n_loc, ld, dtype=torch.bfloat16, device=dev)
    center_split_bf16(feats, mean, hi, lo)
    gram = torch.zeros(d, d, dtype=torch.float32, device=dev)
    nbytes = gemm_tc_workspace_bytes(d, d, n_loc)
    ws = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device=dev)
    for a, b in ((hi, hi), (hi, lo), (lo, hi)):
        gemm_tc(a, 1, ld, b, 1, ld, gram, d, d, n_loc, accumulate=True, workspace=ws)
    if comm is not None:
        comm.allreduce_sum_(gram)
    cov = torch.empty_like(gram)
    rowop(ROW_SCALE_ROWS, gram, torch.full((d,), 1.0 / max(n - 1, 1), dtype=torch.float32, device=dev), cov, d, d)
    return mean, cov


def gemm_tc_workspace_bytes(M, N, K):
    return int(L().jck_gemm_tc_workspace_bytes(M, N, K))


def pack_fc_t(w4, w_t):
    check(L().jck_pack_fc_t(_p(w4), _p(w_t), w4.shape[0], w4.shape[1], _s()), "pack_fc_t")


def unpack_fc_grad_t(dw_t, dw4, accumulate):
    check(L().jck_unpack_fc_grad_t(_p(dw_t), _p(dw4), dw4.shape[0], dw4.shape[1], int(accumulate), _s()), "unpack_fc_grad_t")


def concat_rows(a, b, out):
    """out[m] = [a[m] | b[m] | 0 ...]: a fp32 [M,K1], b fp32 or int64 [M,K2], out fp32 / bf16 [M, ldo >= K1 + K2]"""
    M, K1 = a.shape
    K2 = b.shape[1]
    assert a.dtype == torch.float32 and b.dtype in (torch.float32, torch.int64) and a.is_contiguous() and b.is_contiguous()
    check(L().jck_concat_rows(_p(a), _p(b), int(b.dtype == torch.int64), _p(out), M, K1, K2, out.shape[1], dt(out), _s()),
          "concat_rows")


def zero(t):
    """stream-ordered memset of a contiguous tensor (one memset node in a captured graph)"""
    assert t.is_contiguous()
    check(L().jck_zero(_p(t), t.numel() * t.element_size(), _s()), "zero")


def copy_f32(src, dst):
    assert src.dtype == torch.float32 and dst.dtype == torch.float32 and src.is_contiguous() and dst.is_contiguous()
    check(L().jck_copy_f32(_p(src), _p(dst), src.numel(), _s()), "copy_f32")


class ZeroArena:
    """The fp32 accumulators of one train step (BatchNorm statistics and backward sums of every pass, the logged scalars)
    as slices of ONE buffer zeroed by ONE memset at the start of the step, instead of a torch.zeros fill kernel per pass."""

    def __init__(self, device, numel=1 << 16):
        self.buf = torch.empty(numel, dtype=torch.float32, device=device)
        self.cursor = 0

    def reset(self):
        self.cursor = 0
        zero(self.buf)

    def take(self, n):
        """n zeroed floats, or None when the arena is exhausted (the caller falls back to torch.zeros)"""
        n_pad = -(-n // 32) * 32
        if self.cursor + n_pad > self.buf.numel():
            return None
        out = self.buf[self.cursor:self.cursor + n]
        self.cursor += n_pad
        return out


def cast_rows_bf16(x, out):
    M, K = x.shape
    check(L().jck_cast_rows_bf16(_p(x), _p(out), M, K, out.shape[1], _s()), "cast_rows_bf16")


def gemm_tc(A, a_mn, lda, B, b_mn, ldb, C, M, N, K, accumulate=False, workspace=None, stats=None, stats_channels=0):
    """C[m][n] (+)= sum_k A(m,k) * B(n,k) on tcgen05: bf16 operands, K-major (x_mn = 0: X[row*ldx + k]) or MN-major
    (x_mn = 1: X[k*ldx + row]); C row-major fp32 (accumulate allowed) or bf16, leading dimension C.shape[-1]."""
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    ldc = C.shape[-1] if C.dim() > 1 else N
    wsb = workspace.numel() * workspace.element_size() if workspace is not None else 0
    check(L().jck_gemm_tc(_p(A), int(a_mn), lda, _p(B), int(b_mn), ldb, _p(C), dt(C), ldc, M, N, K, int(accumulate),
                          _p(stats), stats_channels, _p(workspace), wsb, _s()), "gemm_tc")


ROW_BIAS_ACT, ROW_MUL, ROW_ACT_BWD, ROW_ADD_BCAST, ROW_SUM_GROUPS, ROW_OUTER, ROW_SCALE_ROWS, ROW_THRESH = range(8)


def dropout_mask(out, p_drop, seed, stream_id, counter):
    """keep-mask (1 with probability 1 - p_drop) from our Philox stream; advances `counter`."""
    rand(out, seed, stream_id, counter)
    rng_advance(counter, (out.numel() + 3) // 4)
    rowop(ROW_THRESH, out, None, out, out.shape[0], out.shape[1], s=p_drop)
    return out


def rowop(op, x, y, out, M, N, rows_y=0, s=1.0):
    check(L().jck_rowop(op, _p(x), _p(y), _p(out), M, N, rows_y, float(s), _s()), "rowop")


def sigmoid_bce(logit, prob, target, scalars):
    check(L().jck_sigmoid_bce(_p(logit), _p(prob), float(target), _p(scalars), prob.numel(), _s()), "sigmoid_bce")


def logit_grad(prob, out, mode, target=0.0, scale=1.0, up=None):
    check(L().jck_logit_grad(_p(prob), _p(up), float(target), _p(out), prob.numel(), mode, float(scale), _s()), "logit_grad")


def i64_to_f32(x, out):
    check(L().jck_i64_to_f32(_p(x), _p(out), x.numel(), _s()), "i64_to_f32")


def axpy(x, y, a=1.0):
    check(L().jck_axpy(_p(x), _p(y), float(a), x.numel(), dt(x), _s()), "axpy")


def gp_seed(v, u, scalars, scale):
    B = v.shape[0]
    check(L().jck_gp_seed(_p(v), _p(u), _p(scalars), B, v.numel() // B, float(scale), dt(v), _s()), "gp_seed")


def pack_linear(w, w_a, w_b, C, HW):
    O, E = w.shape[0], w.shape[1] - C * HW
    check(L().jck_pack_linear(_p(w), _p(w_a), _p(w_b), O, C, HW, E, dt(w_a), _s()), "pack_linear")


def unpack_linear_grad(dwa, dwb, dw, C, HW, accumulate):
    O, E = dw.shape[0], dw.shape[1] - C * HW
    check(L().jck_unpack_linear_grad(_p(dwa), _p(dwb), _p(dw), O, C, HW, E, int(accumulate), _s()), "unpack_linear_grad")


def bn_adj_reduce(dbar, da, y, ss, mr, sums1, asums, C, count, slope):
    check(L().jck_bn_adj_reduce(_p(dbar), _p(da), _p(y), _p(ss), _p(mr), _p(sums1), _p(asums), y.numel() // C, C,
                                float(count), slope, dt(y), _s()), "bn_adj_reduce")


def bn_adj_apply(dbar, da, y, ss, mr, gamma, sums1, asums, gbar_a, ybar, C, count, slope):
    check(L().jck_bn_adj_apply(_p(dbar), _p(da), _p(y), _p(ss), _p(mr), _p(gamma), _p(sums1), _p(asums), _p(gbar_a),
                               _p(ybar), y.numel() // C, C, float(count), slope, dt(y), _s()), "bn_adj_apply")


def bn_adj_param(asums, mr, dgamma, C, scale=1.0):
    check(L().jck_bn_adj_param(_p(asums), _p(mr), _p(dgamma), C, float(scale), _s()), "bn_adj_param")


# ---- Inception-v3 feature extractor (metrics.py) -----------------------------------------------------------------
def _ints(vals):
    return (_ct.c_int * len(vals))(*[int(v) for v in vals])


def conv_gemm(act, lda, w, scale, bias, out, ldc, geom):
    """jck_conv_gemm (include/jck_b200.h): implicit-GEMM convolution / linear layer on tcgen05; geom is the host int list."""
    g = _ints(geom)
    check(L().jck_conv_gemm(_p(act), lda, _p(w), _p(scale), _p(bias), _p(out), ldc, _ct.cast(g, _ct.c_void_p), len(geom), _s()),
          "conv_gemm")


def im2col(x, in_geom, ldx, patches, B, H, W, C, kh, kw, sy, sx, py, px, Ho, Wo, Kp):
    g = _ints(in_geom)
    check(L().jck_im2col(_p(x), _ct.cast(g, _ct.c_void_p), ldx, _p(patches), B, H, W, C, kh, kw, sy, sx, py, px, Ho, Wo, Kp, _s()),
          "im2col")


def pool3(x, in_geom, ldx, out, out_geom, ldo, B, H, W, C, stride, pad, Ho, Wo, mode, x_lo=0, out_lo=0):
    """x_lo / out_lo: element offsets of the low planes in split-precision mode (include/jck_b200.h), 0 = plain bf16"""
    gi, go = _ints(in_geom), _ints(out_geom)
    if x_lo or out_lo:
        check(L().jck_pool3_split(_p(x), _ct.cast(gi, _ct.c_void_p), ldx, x_lo, _p(out), _ct.cast(go, _ct.c_void_p), ldo, out_lo,
                                  B, H, W, C, stride, pad, Ho, Wo, mode, _s()), "pool3_split")
        return
    check(L().jck_pool3(_p(x), _ct.cast(gi, _ct.c_void_p), ldx, _p(out), _ct.cast(go, _ct.c_void_p), ldo, B, H, W, C, stride, pad,
                        Ho, Wo, mode, _s()), "pool3")


def global_avgpool(x, out_f32, out_bf16, B, HW, C, x_lo=0, out_lo=0):
    if x_lo or out_lo:
        check(L().jck_global_avgpool_split(_p(x), x_lo, _p(out_f32), _p(out_bf16), out_lo, B, HW, C, _s()), "global_avgpool_split")
        return
    check(L().jck_global_avgpool(_p(x), _p(out_f32), _p(out_bf16), B, HW, C, _s()), "global_avgpool")


def resize_norm(x_nchw, out_nhwc, B, C, Hi, Wi, Ho, Wo, ldo, a, b, mean3, std3):
    m, s = (_ct.c_float * 3)(*mean3), (_ct.c_float * 3)(*std3)
    check(L().jck_resize_norm(_p(x_nchw), _p(out_nhwc), B, C, Hi, Wi, Ho, Wo, ldo, a, b, _ct.cast(m, _ct.c_void_p),
                              _ct.cast(s, _ct.c_void_p), _s()), "resize_norm")


def stem_patches(x_nchw, patches, B, Hi, Wi, Hr, Wr, a, b, mean3, std3, patches_lo=0):
    m, s = (_ct.c_float * 3)(*mean3), (_ct.c_float * 3)(*std3)
    if patches_lo:
        check(L().jck_stem_patches_split(_p(x_nchw), _p(patches), patches_lo, B, Hi, Wi, Hr, Wr, a, b, _ct.cast(m, _ct.c_void_p),
                                         _ct.cast(s, _ct.c_void_p), _s()), "stem_patches_split")
        return
    check(L().jck_stem_patches(_p(x_nchw), _p(patches), B, Hi, Wi, Hr, Wr, a, b, _ct.cast(m, _ct.c_void_p),
                               _ct.cast(s, _ct.c_void_p), _s()), "stem_patches")


def inception_score(logits, splits, scores):
    n, d = logits.shape
    check(L().jck_inception_score(_p(logits), n, d, splits, _p(scores), _s()), "inception_score")
