"""ctypes binding of libjck_b200.so (the C ABI declared in include/jck_b200.h).

The product path has no fallback: if the library is missing or a call fails, this raises."""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjck_b200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "jck_b200.h")

JCK_F32, JCK_BF16 = 0, 1
ALGO_AUTO, ALGO_SIMT, ALGO_TC = 0, 1, 2
IMG_NHWC, IMG_P4 = 0, 1

_lib = None

c_p, c_i, c_f, c_ll, c_ull, c_sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong,
                                    ctypes.c_ulonglong, ctypes.c_size_t)

# name -> argtypes (restype is int unless listed in _RESTYPES)
_SIGNATURES = {
    "jck_version": [],
    "jck_last_error_string": [],
    "jck_launch_count": [],
    "jck_prep_image": [c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_nhwc_to_nchw_f32": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_pack_weights_edge": [c_p, c_p, c_p, c_i, c_i, c_p],
    "jck_edge_wgrad_img": [c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_edge_down_img": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_edge_up": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "jck_edge_up_scatter": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "jck_edge_wgrad_workspace_bytes": [c_i, c_i, c_i, c_i],
    "jck_pack_weights": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_conv_down": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_conv_up": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_conv_up_bnbwd": [c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_conv_down_bnbwd": [c_p, c_p, c_p, c_p, c_p, c_f, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_conv_wgrad_workspace_bytes": [c_i, c_i, c_i, c_i, c_i, c_i, c_i],
    "jck_conv_wgrad": [c_p, c_p, c_p, c_p, c_sz, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_dense": [c_p, c_i, c_ll, c_ll, c_p, c_i, c_ll, c_ll, c_p, c_i, c_ll, c_i, c_i, c_i, c_i, c_p],
    "jck_rowop": [c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p],
    "jck_sigmoid_bce": [c_p, c_p, c_f, c_p, c_i, c_p],
    "jck_logit_grad": [c_p, c_p, c_f, c_p, c_i, c_i, c_f, c_p],
    "jck_i64_to_f32": [c_p, c_p, c_ll, c_p],
    "jck_f32_to_bf16": [c_p, c_p, c_ll, c_p],
    "jck_center_split_bf16": [c_p, c_p, c_p, c_p, c_ll, c_i, c_i, c_p],
    "jck_gemm_tc_workspace_bytes": [c_i, c_i, c_i],
    "jck_gemm_tc": [c_p, c_i, c_ll, c_p, c_i, c_ll, c_p, c_i, c_ll, c_i, c_i, c_i, c_i, c_p, c_i, c_p, c_sz, c_p],
    "jck_pack_fc_t": [c_p, c_p, c_i, c_i, c_p],
    "jck_unpack_fc_grad_t": [c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_cast_rows_bf16": [c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_concat_rows": [c_p, c_p, c_i, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_zero": [c_p, c_sz, c_p],
    "jck_copy_f32": [c_p, c_p, c_ll, c_p],
    "jck_axpy": [c_p, c_p, c_f, c_ll, c_i, c_p],
    "jck_gp_seed": [c_p, c_p, c_p, c_i, c_ll, c_f, c_i, c_p],
    "jck_pack_linear": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_unpack_linear_grad": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_bn_adj_reduce": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_f, c_f, c_i, c_p],
    "jck_bn_adj_apply": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_f, c_f, c_i, c_p],
    "jck_bn_adj_param": [c_p, c_p, c_p, c_i, c_f, c_p],
    "jck_fc_fwd": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_fc_wgrad": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_pack_fc": [c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_unpack_fc_grad": [c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_bn_finalize": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_f, c_f, c_p],
    "jck_bn_act_fwd": [c_p, c_p, c_p, c_ll, c_i, c_ll, c_f, c_i, c_i, c_p],
    "jck_bn_act_bwd_reduce": [c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_ll, c_f, c_i, c_i, c_p],
    "jck_bn_act_bwd_apply": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_ll, c_f, c_f, c_i, c_i, c_p],
    "jck_bn_param_grad": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_head_fwd": [c_p, c_p, c_p, c_f, c_p, c_i, c_i, c_i, c_p],
    "jck_head_bwd": [c_p, c_p, c_f, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_pack_head": [c_p, c_p, c_i, c_i, c_p],
    "jck_unpack_head_grad": [c_p, c_p, c_i, c_i, c_p],
    "jck_prep_image_rng": [c_p, c_ull, c_ull, c_p, c_f, c_f, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_g_out_fwd_rng": [c_p, c_ull, c_ull, c_p, c_f, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_g_out_fwd": [c_p, c_p, c_f, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_g_out_bwd": [c_p, c_p, c_f, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_gp_penalty": [c_p, c_p, c_i, c_ll, c_i, c_p],
    "jck_adam": [c_p, c_p, c_p, c_p, c_ll, c_f, c_f, c_f, c_f, c_p, c_p],
    "jck_adam_advance": [c_p, c_p],
    "jck_randn": [c_p, c_ll, c_ull, c_ull, c_p, c_p],
    "jck_rand": [c_p, c_ll, c_ull, c_ull, c_p, c_p],
    "jck_rng_advance": [c_p, c_ull, c_p],
    "jck_conv_gemm": [c_p, c_ll, c_p, c_p, c_p, c_p, c_ll, c_p, c_i, c_p],
    "jck_im2col": [c_p, c_p, c_ll, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_pool3": [c_p, c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_global_avgpool": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "jck_resize_norm": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p],
    "jck_stem_patches": [c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p],
    "jck_pool3_split": [c_p, c_p, c_ll, c_ll, c_p, c_p, c_ll, c_ll, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_p],
    "jck_global_avgpool_split": [c_p, c_ll, c_p, c_p, c_ll, c_i, c_i, c_i, c_p],
    "jck_stem_patches_split": [c_p, c_p, c_ll, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p],
    "jck_inception_score": [c_p, c_i, c_i, c_i, c_p, c_p],
    "jck_u8_resize_norm": [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_p, c_p, c_p],
    "jck_one_hot_i64": [c_p, c_p, c_p, c_i, c_i, c_p],
    "jck_comm_create": [c_i, c_i, ctypes.POINTER(c_p), c_p],
    "jck_comm_connect": [c_p, c_p],
    "jck_comm_destroy": [c_p],
    "jck_comm_configure": [c_p, c_i],
    "jck_comm_error": [c_p, ctypes.POINTER(c_i)],
    "jck_comm_allreduce_small": [c_p, c_p, c_i, c_p],
    "jck_bn_finalize_sync": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_f, c_f, c_p],
    "jck_bn_bwd_sums_sync": [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
}
_RESTYPES = {"jck_last_error_string": ctypes.c_char_p, "jck_launch_count": c_ull,
             "jck_conv_wgrad_workspace_bytes": c_sz, "jck_edge_wgrad_workspace_bytes": c_sz,
             "jck_gemm_tc_workspace_bytes": c_sz}


class JckError(RuntimeError):
    pass


def header_symbols():
    """Every function name include/jck_b200.h declares."""
    with open(HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    return sorted(set(re.findall(r"\b(jck_[a-z0-9_]+)\s*\(", text)))


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JckError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a). "
                       "There is no CPU or PyTorch fallback for the train step.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_i)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().jck_last_error_string()
        raise JckError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(load().jck_launch_count())
