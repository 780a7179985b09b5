"""ctypes binding of libmi_b200.so (the C ABI declared in include/mi_b200.h).

The product path has no CPU fallback: if the library is missing or a kernel
returns an error this raises, it never reroutes to another implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmi_b200.so")

_f = C.c_void_p      # const float* / float*  (raw device addresses)
_i = C.c_int
_fl = C.c_float
_sz = C.c_size_t
_st = C.c_void_p     # cudaStream_t

# name -> (restype, argtypes); mirrors include/mi_b200.h one to one
SIGNATURES = {
    "mi_version": (_i, []),
    "mi_error_string": (C.c_char_p, [_i]),
    "mi_launch_count": (C.c_ulonglong, []),
    "mi_tc_available": (_i, []),
    "mi_set_sm_budget": (_i, [_i]),
    "mi_wgrad_defer_begin": (_i, []),
    "mi_wgrad_defer_flush": (_i, [_st]),
    "mi_set_pad_lanes_scratch": (_i, [_i]),
    "mi_prof_enable": (_i, [_i]),
    "mi_prof_summary": (_i, [_i, C.c_void_p]),
    "mi_conv2d_fprop": (_i, [_f, _i, _f, _i, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _fl, _i, _st]),
    "mi_conv2d_dgrad": (_i, [_f, _i, _f, _i, _f, _i, _f, _i, _i, _fl, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_weight_to_dgrad": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _st]),
    "mi_round_tf32": (_i, [_f, _i, _f, _i, _i, _sz, _st]),
    "mi_conv2d_wgrad_workspace": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "mi_conv2d_wgrad": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _fl, _f, _f, _f, _f, _f, _f, _f, _f, _f,
                             _f, _f, _i, _f, _f, _sz, _i, _st]),
    "mi_avgpool2_fwd": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _st]),
    "mi_avgpool2_bwd": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _st]),
    "mi_maxpool2_fwd": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _st]),
    "mi_maxpool2_bwd": (_i, [_f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _i, _st]),
    "mi_upsample2_fwd": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_upsample2_bwd": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_upsample2_window_fwd": (_i, [_f, _i, _f, _i] + [_i] * 14 + [_st]),
    "mi_upsample2_window_bwd": (_i, [_f, _i, _f, _i] + [_i] * 14 + [_f, _i, _i, _fl, _i, _st]),
    "mi_window_copy": (_i, [_f, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_add": (_i, [_f, _i, _f, _i, _f, _i, _sz, _i, _i, _st]),
    "mi_copy": (_i, [_f, _i, _f, _i, _i, _sz, _i, _st]),
    "mi_act_bwd": (_i, [_f, _i, _f, _i, _i, _fl, _sz, _i, _i, _st]),
    "mi_fill": (_i, [_f, _fl, _sz, _st]),
    "mi_bn_eval_fwd": (_i, [_f, _i, _f, _i, _f, _f, _f, _f, _fl, _i, _fl, _sz, _i, _st]),
    "mi_bn_eval_bwd_workspace": (_sz, [_sz, _i]),
    "mi_bn_eval_bwd": (_i, [_f, _i, _f, _i, _f, _i, _f, _i, _i, _f, _f, _f, _fl, _i, _fl, _f, _f, _i, _fl, _f, _sz,
                            _sz, _i, _st]),
    "mi_binary_fwd": (_i, [_i, _f, _i, _f, _i, _i, _f, _i, _sz, _i, _st]),
    "mi_binary_bwd": (_i, [_i, _f, _i, _f, _i, _i, _f, _i, _f, _i, _i, _f, _i, _i, _sz, _i, _st]),
    "mi_affine": (_i, [_f, _i, _f, _i, _fl, _fl, _i, _sz, _i, _st]),
    "mi_act_fwd": (_i, [_f, _i, _f, _i, _i, _fl, _sz, _i, _st]),
    "mi_clamp_fwd": (_i, [_f, _i, _f, _i, _fl, _fl, _sz, _i, _st]),
    "mi_clamp_bwd": (_i, [_f, _i, _f, _i, _f, _i, _i, _fl, _fl, _sz, _i, _st]),
    "mi_blend_fwd": (_i, [_f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _fl, _fl, _fl, _i, _sz, _i, _st]),
    "mi_blend_bwd": (_i, [_f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _i, _fl, _fl, _fl,
                          _i, _sz, _i, _st]),
    "mi_ring_fix": (_i, [_f, _i, _i, _i, _i, _i, _i, _st]),
    "mi_ring_fold": (_i, [_f, _i, _i, _i, _i, _i, _i, _st]),
    "mi_channel_mean_nchw": (_i, [_f, _f, _i, _i, _st]),
    "mi_space_to_depth": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_depth_to_space": (_i, [_f, _i, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_depth_to_space_bwd": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_interior_reduce": (_i, [_f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _fl, _st]),
    "mi_scale_add": (_i, [_f, _i, _f, _f, _i, _f, _i, _i, _sz, _i, _st]),
    "mi_scale_bwd": (_i, [_f, _i, _f, _f, _i, _i, _i, _sz, _i, _st]),
    "mi_interior_bcast_add": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _fl, _st]),
    "mi_frames_to_canvas": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_nhwc_window_to_nchw": (_i, [_f, _i, _f, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_nchw_to_nhwc_window": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _i, _st]),
    "mi_sepconv_planar_bytes": (_sz, [_i, _i, _i, _i]),
    "mi_sepconv_fwd": (_i, [_f, _f, _f, _i, _f] + [_i] * 13 + [_f, _st]),
    "mi_sepconv_bwd": (_i, [_f, _f, _f, _i, _f, _f, _f, _i] + [_i] * 14 + [_f, _i, _f, _st]),
    "mi_warp_fwd": (_i, [_f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _i, _fl, _fl, _st]),
    "mi_warp_bwd": (_i, [_f, _i, _f, _i, _f, _i, _f, _i, _f, _i, _i, _i, _i, _i, _i, _i, _fl, _fl, _st]),
    "mi_loss_fwd_bwd": (_i, [_f, _f, _f, _f, _sz, _i, _fl, _st]),
    "mi_psnr_accumulate": (_i, [_f, _f, _f, _sz, _st]),
    "mi_ssim_accumulate": (_i, [_f, _f, _f, _i, _i, _i, C.c_void_p, _i, _fl, _st]),
    "mi_inner_update": (_i, [_f, _f, _f, _f, _f, _f, _i, _i, _i, _f, _f, _sz, _i, _i, _st]),
    "mi_outer_step": (_i, [_f, _f, _f, _f, _sz, _i, _fl, _fl, _fl, _fl, _fl, _i, _st]),
    "mi_axpby": (_i, [_f, _fl, _f, _fl, _sz, _st]),
    "mi_addcmul": (_i, [_f, _fl, _f, _f, _sz, _st]),
    "mi_segment_dot": (_i, [_f, _f, _f, _f, _sz, _st]),
    "mi_segment_scale": (_i, [_f, _f, _f, _f, _f, _fl, _i, _sz, _st]),
    # (src, dst, y0, x0, reversed: device addresses; mean3 / std3: HOST float[3] or NULL)
    "mi_septuplet_prepare": (_i, [_f, _f, _f, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, C.c_void_p, C.c_void_p, _st]),
}


class MiB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load libmi_b200.so and bind every entry point; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise MiB200Error(
            "libmi_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C meta_interpolation_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code, what=""):
    if code != 0:
        msg = load().mi_error_string(int(code))
        raise MiB200Error("%s failed: %s (code %d)" % (what or "libmi_b200 call", msg.decode() if msg else "?", code))
