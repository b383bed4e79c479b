"""ctypes binding of libdad_b200.so (include/dad_b200.h).  There is no fallback: if the library
is missing or a call fails, an exception is raised."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libdad_b200.so")

_c = ctypes
_vp, _i, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t


class ModelDesc(ctypes.Structure):
    _fields_ = [("embed_dim", _i), ("depth", _i), ("num_heads", _i), ("taps", _i * 4),
                ("features", _i), ("out_channels", _i * 4)]


# name -> (restype, argtypes); mirrors include/dad_b200.h one to one
PROTOTYPES = {
    "dad_last_error": (_c.c_char_p, []),
    "dad_abi_version": (_i, []),
    "dad_model_create": (_i, [_c.POINTER(ModelDesc), _c.POINTER(_vp)]),
    "dad_model_destroy": (None, [_vp]),
    "dad_model_set_weight": (_i, [_vp, _c.c_char_p, _vp, _i64, _vp]),
    "dad_model_prepare": (_i, [_vp, _i, _i, _i, _vp]),
    "dad_forward_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "dad_forward": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dad_model_debug_capture": (_i, [_vp, _c.c_char_p, _vp, _i64]),
    "dad_model_set_grad": (_i, [_vp, _c.c_char_p, _vp, _i64]),
    "dad_train_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "dad_forward_train": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dad_backward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dad_loss_workspace_bytes": (_sz, [_i, _i]),
    "dad_masked_shift_and_scale": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_ssi_loss": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dad_contexts_dr": (_i, [_i, _vp, _vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "dad_contexts_dp": (_i, [_i, _vp, _vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "dad_contexts_ds": (_i, [_i, _vp, _i, _i, _i, _vp, _vp]),
    "dad_ssi_loss_bwd": (_i, [_vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_hdn_loss_dr_bwd": (_i, [_i, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_hdn_loss_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_grad_loss_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "dad_feat_cos_loss_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "dad_distill_loss_bwd": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_preprocess_image": (_i, [_vp, _i, _i, _i64, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "dad_resize_depth": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "dad_minmax_normalize": (_i, [_vp, _i, _i64, _vp, _vp, _sz, _vp]),
    "dad_colorize_depth": (_i, [_vp, _vp, _i, _i64, _c.c_float, _c.c_float, _i, _vp, _vp, _vp, _vp, _vp]),
    "dad_hdn_loss_dr": (_i, [_i, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_ssi_hdn_dr_loss": (_i, [_i, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dad_hdn_loss": (_i, [_vp, _vp, _vp, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dad_grad_loss": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dad_feat_cos_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dad_distill_loss": (_i, [_vp, _vp, _i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dad_gemm": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "dad_gemm_ex": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dad_gemm_splitk": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "dad_conv_nhwc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dad_conv_nhwc_ex": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dad_attention": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "dad_gemm_splitk_mn": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dad_conv_wgrad": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "dad_gemm_shifted": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _c.POINTER(_c.c_int), _i, _vp]),
    "dad_layernorm": (_i, [_vp, _vp, _vp, _vp, _vp, _c.c_longlong, _i, _i, _i, _i, _c.c_float, _i, _vp]),
    "dad_launch_count": (_c.c_longlong, []),
    "dad_profile_enable": (None, [_i]),
    "dad_profile_get": (_i, [_i, _c.POINTER(_c.c_double), _c.POINTER(_c.c_double), _c.POINTER(_c.c_longlong)]),
}

DAD_ERR_INVALID, DAD_ERR_UNSUPPORTED, DAD_ERR_CUDA, DAD_ERR_WORKSPACE = -1, -2, -3, -4
_lib = None


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU / eager fallback for this path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=""):
    if code == 0:
        return
    msg = load().dad_last_error().decode("utf-8", "replace")
    text = f"{what}: {msg}" if what else msg
    if code == DAD_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    if code == DAD_ERR_INVALID:
        raise ValueError(text)
    raise RuntimeError(f"{text} (code {code})")


def ptr(t):
    """Device pointer of a tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
