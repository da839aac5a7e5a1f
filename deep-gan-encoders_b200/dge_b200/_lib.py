"""ctypes binding of libdge_b200.so -- the C ABI declared in include/dge_b200.h.

The library is the product path: if it is missing or the device is not a B200 the ops raise
(there is NO CPU / eager fallback anywhere in this package).
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DGE_LIB_PATH", os.path.join(_HERE, "libdge_b200.so"))  # env override: experiments only

_lib = None


class DgeError(RuntimeError):
    pass


class ConvArgs(Structure):
    """Mirror of `dge_conv_args` (include/dge_b200.h)."""
    _fields_ = [
        ("kind", c_int32), ("flags", c_int32),
        ("n", c_int32), ("h", c_int32), ("w", c_int32),
        ("cin", c_int32), ("cout", c_int32),
        ("planes", c_int32),
        ("x", c_void_p), ("wpk", c_void_p),
        ("demod", c_void_p), ("noise", c_void_p), ("noise_bstride", c_int64),
        ("noise_w", c_void_p), ("noise_scalar", c_float),
        ("bias", c_void_p), ("slope", c_float), ("gain", c_float),
        ("blend_src", c_void_p), ("blend_pool", c_int32), ("blend_a", c_float), ("blend_b", c_float),
        ("out_act", c_void_p), ("out_planes", c_int32), ("out_scale", c_void_p),
        ("out_f32b", c_void_p), ("out_nchw", c_void_p),
        ("rgb_w", c_void_p), ("rgb_out", c_void_p),
        ("out_raw_up", c_void_p),
        ("preact_add", c_void_p),
        ("preact_c", c_int32), ("preact_up", c_int32),
        ("splitk_ws", c_void_p),
        ("out_pool", c_int32),
        ("in_h", c_int32), ("in_w", c_int32),
    ]


class Sg2PrepItem(Structure):
    """Mirror of `dge_sg2_prep_item` (include/dge_b200.h)."""
    _fields_ = [
        ("st_w", c_void_p), ("st_b", c_void_p), ("w2", c_void_p), ("rgb_w", c_void_p),
        ("style_off", c_int64), ("demod_off", c_int64), ("rgbw_off", c_int64),
        ("wp_index", c_int32), ("cin", c_int32), ("cout", c_int32), ("nch", c_int32),
        ("st_wscale", c_float), ("st_bscale", c_float), ("st_add_bias", c_float), ("rgb_scale", c_float),
        ("eps", c_float), ("pad_", c_int32),
    ]


# name -> (restype, argtypes); every symbol include/dge_b200.h declares
P = c_void_p
SIGNATURES = {
    "dge_last_error": (c_char_p, []),
    "dge_version": (c_int, []),
    "dge_device_ok": (c_int, []),
    "dge_launch_count": (c_int64, []),
    "dge_launch_count_reset": (None, []),
    "dge_conv_forward": (c_int, [POINTER(ConvArgs), P]),
    "dge_conv_splitk_ws_bytes": (ctypes.c_size_t, [POINTER(ConvArgs)]),
    "dge_pack_conv_weight": (c_int, [P, P, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    "dge_pack_conv_weight_dgrad": (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, P]),
    "dge_conv_wgrad": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_weight_sqsum": (c_int, [P, P, c_int, c_int, c_int, c_float, P]),
    "dge_demod": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P]),
    "dge_rgb_weights": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P]),
    "dge_sg2_prep": (c_int, [P, c_int, P, P, c_int, c_int, c_int, P]),
    "dge_sg2_prep_bwd": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, P]),
    "dge_dense": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_float, P]),
    "dge_pixel_norm": (c_int, [P, P, c_int, c_int, c_float, P]),
    "dge_nchw_to_act": (c_int, [P, c_int64, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_nchw_to_f32b": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "dge_f32b_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "dge_act_to_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_f32b_to_act": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_to_rgb_nchw": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_up_fir_epilogue": (c_int, [P, P, P, c_int64, c_float, P, c_float, c_float, P, P, P,
                                    c_int, c_int, c_int, c_int, c_int, P]),
    "dge_rgb_init": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "dge_from_rgb": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "dge_instance_stats": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "dge_instance_norm": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_instance_norm_pool": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_from_rgb_stats": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, P]),
    "dge_avgpool_to_act": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_lreq_adam_step": (c_int, [P, P, P, P, P, P, P, c_int, c_int, c_float, c_float, P]),
    "dge_pair_moments": (c_int, [P, P, c_int64, P, P]),
    "dge_softmax_kl_sum": (c_int, [P, P, c_int64, c_int, c_int64, P, P]),
    "dge_avgpool_nchw": (c_int, [P, P, c_int64, c_int, c_int, c_int, P]),
    "dge_ssim_sum": (c_int, [P, P, c_int64, c_int, c_int, P, P]),
    "dge_ssim_grad": (c_int, [P, P, P, P, P, c_int64, c_int, c_int, P]),
    "dge_instance_norm_affine": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_pixelnorm_to_act": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_float, c_int, P]),
    "dge_pixelnorm_to_rgb": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P]),
    "dge_upsample_nearest_nchw": (c_int, [P, P, c_int64, c_int, c_int, P]),
    "dge_axpby": (c_int, [P, P, P, c_float, c_float, c_int64, P]),
    "dge_sg1_post": (c_int, [P, c_int, P, P, P, c_float, P, c_int, c_int, c_int, c_int, P]),
    "dge_instance_norm_style": (c_int, [P, c_int, P, P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_to_rgb_f32b": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_instance_norm_blur": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_cbn_coeffs": (c_int, [P, P, P, P, P, P, c_float, P, P, c_int, c_int, P]),
    "dge_affine_act": (c_int, [P, P, P, c_int, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_maxpool2_f32b": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "dge_channel_softmax_to_act": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_tanh_slice_nchw": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "dge_argmax_mode": (c_int, [P, c_int, c_int, P, P, P]),
    "dge_gradcam": (c_int, [P, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_mask2cam": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P]),
    "dge_be_head_bwd": (c_int, [P, P, P, c_float, c_float, c_float, P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_in_bwd_stats": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P]),
    "dge_in_bwd_apply": (c_int, [P, P, P, P, P, P, P, c_int, P, c_float, c_int, P, c_float, P, P, P, c_int, c_int, c_int,
                                 c_int, c_int, P]),
    "dge_affine_relu_bwd": (c_int, [P, P, P, P, c_float, c_int, P, c_int, c_int, P, P, P, c_int, c_int, c_int, c_int,
                                    c_int, P]),
    "dge_from_rgb_bwd": (c_int, [P, P, P, P, c_float, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_sg2_layer_bwd": (c_int, [P, P, P, P, P, P, c_int64, c_float, P, P, c_float, c_float, P, c_int, P, P, c_int, c_int,
                                  c_int, c_int, c_int, P]),
    "dge_up_fir_bwd_s2d": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_rgb_up_bwd": (c_int, [P, P, c_int64, c_int, c_int, P]),
    "dge_lpips_input": (c_int, [P, P, c_float, c_float, c_float, c_float, c_float, c_float, c_int, c_int, c_int, c_int, P]),
    "dge_maxpool_to_act": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_relu_pool_bwd": (c_int, [P, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "dge_lpips_dist": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_float, P]),
    "dge_blend": (c_int, [P, P, P, c_float, c_float, c_int, c_int, c_int, c_int, c_int, P]),
}


def load():
    """Load the shared library (once). Raises DgeError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DgeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C deep-gan-encoders_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dge_last_error()
        raise DgeError(f"dge_b200 call failed (rc={rc}): {msg.decode() if msg else '?'}")
