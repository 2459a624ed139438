"""ctypes binding of the C ABI in include/mpn_b200.h (libmpn_b200.so, built in-tree by csrc/build.py).

No fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPN_B200_LIB: an instrumented build of the same ABI (csrc/build.py --trace -> libmpn_b200_trace.so, scripts/exp/trace_conv.py)
LIB_PATH = os.environ.get("MPN_B200_LIB") or os.path.join(_HERE, "csrc", "libmpn_b200.so")

FMT_F32, FMT_BF16, FMT_BF16X2, FMT_F16F8 = 0, 1, 2, 3
OUT_ACT, OUT_F32_NHWC, OUT_F32_NCHW = 0, 1, 2
EPI_RELU, EPI_SIGMOID, EPI_NO_H8, IN_NO_H8, IN_DERIVE_H8, W_MERGED = 1, 2, 4, 8, 16, 32

c_void_p, c_int, c_float, c_size_t, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in (
        "N", "H", "W", "Cin", "Cout", "R", "S", "stride", "pad", "OH", "OW", "fmt", "in_cstride", "flags",
        "res_cstride", "up_h", "up_w", "up_cstride", "out_mode", "out_cstride", "out_coffset", "out_rep")] + [
        ("out_nstride", c_ll), ("w_cout_pad", c_int), ("in_wpitch", c_int), ("in_hpitch", c_int), ("k_overlap", c_int),
        ("acc_scale", c_float), ("gat_n", c_int), ("gat_shift", c_int * 2), ("gat_h", c_int * 2), ("gat_w", c_int * 2),
        ("gat_cstride", c_int * 2)]


class ConvPtrs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in (
        "x_hi", "x_lo", "w_hi", "w_lo", "scale", "bias", "res_hi", "res_lo", "up_hi", "up_lo", "y_hi", "y_lo")] + [
        ("gat_hi", c_void_p * 2), ("gat_lo", c_void_p * 2)]


class MpnError(RuntimeError):
    pass


_lib = None

_SIGS = {
    "mpn_last_error": (ctypes.c_char_p, []),
    "mpn_version": (c_int, []),
    "mpn_sizeof_conv_desc": (c_int, []),
    "mpn_sizeof_conv_ptrs": (c_int, []),
    "mpn_device_supports_tcgen05": (c_int, []),
    "mpn_conv2d_fwd": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvPtrs), c_void_p]),
    "mpn_conv2d_fwd_multi": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvPtrs), c_int, c_void_p]),
    "mpn_conv2d_fwd_f32in": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvPtrs), c_void_p]),
    "mpn_pack_filter_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_pack_filter_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_pack_filter_bf16_scaled": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_filter_absmax": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mpn_pack_filter_f16f8": (c_int, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_fold_bn": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_void_p]),
    "mpn_stem_pack_input": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_stem_pack_filter": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "mpn_preprocess_u8_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mpn_stem_pack_input_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_nchw_to_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_nhwc_to_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_maxpool3x3s2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_relu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "mpn_conv2d_wgrad": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(ConvPtrs), c_void_p, c_void_p]),
    "mpn_pack_filter_dgrad_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_unpack_filter_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_stem_unpack_filter_grad": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "mpn_zero_insert2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_channel_sums": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mpn_double_to_float": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "mpn_bn_stats": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mpn_bn_update_running": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_float, c_int, c_void_p]),
    "mpn_bn_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                             c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "mpn_bn_backward": (c_int, [c_void_p] * 6 + [c_void_p, c_void_p, c_void_p, c_float, c_int, c_ll, c_int, c_int] + [c_void_p] * 9),
    "mpn_relu_backward": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_void_p]),
    "mpn_add_act": (c_int, [c_void_p] * 6 + [c_ll, c_int, c_void_p]),
    "mpn_maxpool3x3s2_backward": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_block_sum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_mse_heatmap_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_float, c_void_p]),
    "mpn_add_softmax_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_num_anchors": (c_int, [c_int, c_int]),
    "mpn_generate_anchors": (c_int, [c_int, c_int, c_void_p]),
    "mpn_decode_clip": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "mpn_detect_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mpn_filter_sort_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpn_filter_sort_nms_profile": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_int, c_int, c_void_p, c_void_p,
                                            c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mpn_nms_workspace_bytes": (c_size_t, [c_int]),
    "mpn_nms": (c_int, [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpn_nms_mask": (c_int, [c_void_p, c_int, c_float, c_int, c_void_p, c_void_p]),
    "mpn_heatmap_peaks_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mpn_heatmap_peaks": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_ll, c_float, c_int, c_void_p, c_int, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "mpn_resize_cubic_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mpn_resize_cubic": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int, ctypes.c_double,
                                 ctypes.c_double, c_int, c_float, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mpn_tta_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll, c_void_p]),
    "mpn_focal_loss_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mpn_focal_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_float, c_float, c_void_p, c_size_t, c_void_p]),
    "mpn_prn_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mpn_prn_build_inputs": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, ctypes.c_double,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "mpn_prn_assign": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
}

EXPORTS = tuple(_SIGS.keys())


def lib():
    """The loaded library; raises MpnError if it has not been built (python __graft_entry__.py / csrc/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MpnError("libmpn_b200.so is missing (%s). Build it with `python -m multiposenet.pytorch_b200.csrc.build` "
                           "or __graft_entry__.build(); there is no CPU / eager fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        for struct, size in ((ConvDesc, l.mpn_sizeof_conv_desc()), (ConvPtrs, l.mpn_sizeof_conv_ptrs())):
            if ctypes.sizeof(struct) != size:
                raise MpnError("%s is %d bytes here but %d in libmpn_b200.so: rebuild the library (include/mpn_b200.h changed)"
                               % (struct.__name__, ctypes.sizeof(struct), size))
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mpn_last_error()
        raise MpnError("%s failed (%d): %s" % (what or "mpn call", rc, msg.decode() if msg else "?"))
