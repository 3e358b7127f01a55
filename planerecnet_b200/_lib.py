"""ctypes binding of libprn_b200.so (the C ABI declared in include/prn_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libprn_b200.so")

PRN_F16, PRN_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS, ACT_DCN_OFFMASK, ACT_SIGMOID_AVG4 = range(6)
PAD_ZERO, PAD_REFLECT, PAD_CLAMP = 0, 1, 2


class PrnError(RuntimeError):
    pass


class PrnConv(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p),
        ("c0", C.c_int32), ("c1", C.c_int32),
        ("ld0", C.c_int32), ("ld1", C.c_int32),
        ("batch", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("upsample", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("pad_mode", C.c_int32),
        ("h_out", C.c_int32), ("w_out", C.c_int32),
        ("dcn_offmask", C.c_void_p),
        ("weight", C.c_void_p),
        ("n_pad", C.c_int32), ("w_rows_total", C.c_int32), ("w_group_rows", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ld_res", C.c_int32),
        ("act", C.c_int32), ("act_param", C.c_float),
        ("out16", C.c_void_p), ("ld_out16", C.c_int32),
        ("out32", C.c_void_p), ("ld_out32", C.c_int32),
        ("out_img_rows", C.c_int32),
        ("stats", C.c_void_p), ("stats_cg", C.c_int32),
        ("dtype", C.c_int32),
        ("shuffle_n", C.c_int32),
    ]


class PrnWgrad(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p),
        ("c0", C.c_int32), ("c1", C.c_int32),
        ("ld0", C.c_int32), ("ld1", C.c_int32),
        ("batch", C.c_int32), ("h_in", C.c_int32), ("w_in", C.c_int32),
        ("upsample", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("pad_mode", C.c_int32),
        ("h_out", C.c_int32), ("w_out", C.c_int32),
        ("dy", C.c_void_p), ("n", C.c_int32), ("ld_dy", C.c_int32),
        ("dw", C.c_void_p), ("ld_dw", C.c_int32),
        ("dtype", C.c_int32), ("flags", C.c_int32),
    ]


_lib = None


def lib():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PrnError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the PlaneRecNet hot path)")
    l = C.CDLL(LIB_PATH)
    if not hasattr(l, "prn_build_fingerprint"):
        raise PrnError(f"{LIB_PATH} predates the source tree (no build fingerprint): rebuild it with __graft_entry__.build()")
    l.prn_build_fingerprint.restype = C.c_char_p
    from .csrc import build as _B
    built, want = l.prn_build_fingerprint().decode(), _B.fingerprint()
    if built != want:
        raise PrnError(f"{LIB_PATH} was built from other sources (fingerprint {built[:12]} != {want[:12]}): "
                       "rebuild it with `python -c 'import __graft_entry__ as g; g.build()'`")
    l.prn_last_error.restype = C.c_char_p
    l.prn_abi_version.restype = C.c_int
    l.prn_device_sm_count.restype = C.c_int
    for name in EXPORTS:
        if not hasattr(l, name):
            raise PrnError(f"{LIB_PATH} does not export {name}")
        getattr(l, name).restype = C.c_char_p if name in ("prn_last_error", "prn_build_fingerprint") else C.c_int
    _lib = l
    return l


# every symbol include/prn_b200.h declares (tests check the header against this list and the .so)
EXPORTS = [
    "prn_last_error", "prn_abi_version", "prn_build_fingerprint", "prn_device_sm_count",
    "prn_conv2d_fwd", "prn_conv2d_fwd_profile", "prn_conv2d_plan", "prn_conv2d_plan_ex",
    "prn_stem_im2col", "prn_stem_im2col_image", "prn_maxpool3x3s2", "prn_avgpool2x2", "prn_resize_bilinear", "prn_append_coord",
    "prn_groupnorm_apply", "prn_upsample2x_bilinear", "prn_mul", "prn_ppa_gather",
    "prn_nhwc_to_nchw_f32", "prn_nchw_f32_to_nhwc",
    "prn_conv3x3_to1_reflect", "prn_conv3x3_to1_reflect_devbias", "prn_point_nms_sigmoid", "prn_mask_stats", "prn_upsample_mask_box", "prn_pack_mask_bits", "prn_mask_nms_greedy", "prn_numpy_choice_shuffle",
    "prn_conv2d_wgrad", "prn_conv2d_wgrad_plan", "prn_bn_finalize", "prn_bn_apply", "prn_bn_finalize_apply", "prn_chan_reduce", "prn_bn_bwd_apply",
    "prn_relu_bwd", "prn_add_strided", "prn_add_f32", "prn_add16", "prn_maxpool3x3s2_bwd", "prn_dcn_im2col",
    "prn_dcn_col2im_bwd", "prn_gn_bwd_reduce", "prn_gn_bwd_apply", "prn_avgpool2x2_bwd", "prn_upsample2x_bilinear_bwd",
    "prn_resize_bilinear_bwd", "prn_reflect_fold", "prn_softplus_bwd_pad", "prn_pack_conv_weight", "prn_pack_dgrad_weight", "prn_pack_multi", "prn_copy_multi_f32", "prn_unpack_wgrad_multi", "prn_adam_multi", "prn_adam_multi_checked",
    "prn_focal_loss", "prn_depth_rmselog_fwd", "prn_depth_rmselog_bwd", "prn_dice_lava_rows", "prn_dice_lava_bwd",
    "prn_lava_weights", "prn_vnl_triplets_fwd", "prn_vnl_triplets_bwd",
]


def check(rc, what=""):
    if rc != 0:
        msg = lib().prn_last_error().decode("utf-8", "replace")
        raise PrnError(f"{what or 'libprn_b200'} failed (status {rc}): {msg}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
