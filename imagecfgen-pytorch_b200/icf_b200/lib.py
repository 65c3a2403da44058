"""ctypes binding of libicf_b200.so (the C-ABI declared in include/icf.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libicf_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_LRELU, ACT_TANH = 0, 1, 2
FORM_GATHER, FORM_TRANSPOSED = 0, 1
MAX_PLANES = 8

_vp = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f32 = C.c_float


class ConvArgs(C.Structure):
    _fields_ = [("dtype", _i32), ("form", _i32), ("N", _i32),
                ("H", _i32), ("W", _i32), ("C", _i32), ("in_pitch", _i32),
                ("P", _i32), ("Q", _i32), ("K", _i32), ("out_pitch", _i32),
                ("R", _i32), ("S", _i32), ("stride", _i32), ("pad", _i32),
                ("w_rows", _i32), ("w_pitch", _i32), ("act", _i32), ("slope", _f32), ("out_f32", _i32),
                ("mask_pitch", _i32), ("accumulate", _i32), ("win", _i32),
                ("src", _vp), ("w", _vp), ("bias", _vp), ("dst", _vp), ("out_mask", _vp), ("stats", _vp)]


class WgradArgs(C.Structure):
    _fields_ = [("dtype", _i32), ("N", _i32),
                ("P", _i32), ("Q", _i32), ("A", _i32), ("a_pitch", _i32),
                ("H", _i32), ("W", _i32), ("B", _i32), ("b_pitch", _i32),
                ("R", _i32), ("S", _i32), ("stride", _i32), ("pad", _i32),
                ("small_t", _vp), ("big_t", _vp), ("dw", _vp), ("win", _i32)]


class Perm(C.Structure):
    _fields_ = [("d0", _i64), ("d1", _i64), ("d2", _i64), ("s0", _i64), ("s1", _i64), ("s2", _i64),
                ("d2_pad", _i64), ("d0_pad", _i64)]


class Perm4(C.Structure):
    _fields_ = [("d0", _i64), ("d1", _i64), ("d2", _i64), ("d3", _i64), ("s0", _i64), ("s1", _i64), ("s2", _i64),
                ("s3", _i64), ("d3_pad", _i64), ("row_pitch", _i64)]


class PackJob(C.Structure):
    _fields_ = [("src", _vp), ("dst", _vp), ("dst_dtype", _i32), ("kind", _i32), ("p", Perm), ("p4", Perm4)]


class ImgFeatArgs(C.Structure):
    _fields_ = [("dtype", _i32), ("N", _i32), ("H", _i32), ("W", _i32), ("feat_pitch", _i32),
                ("x_dtype", _i32), ("x_pitch", _i32), ("n_emb", _i32), ("n_cont", _i32),
                ("mask_pitch", _i32), ("pad", _i32),
                ("x", _vp), ("emb_table", _vp * MAX_PLANES), ("emb_index", _vp * MAX_PLANES),
                ("cont", _vp * MAX_PLANES), ("mask", _vp), ("feat", _vp), ("dfeat", _vp),
                ("demb_table", _vp * MAX_PLANES)]


class LatFeatArgs(C.Structure):
    _fields_ = [("dtype", _i32), ("N", _i32), ("latent", _i32), ("feat_pitch", _i32),
                ("z_dtype", _i32), ("z_pitch", _i32), ("n_emb", _i32), ("n_cont", _i32),
                ("emb_k", _i32 * MAX_PLANES),
                ("z", _vp), ("emb_table", _vp * MAX_PLANES), ("onehot", _vp * MAX_PLANES),
                ("cont", _vp * MAX_PLANES), ("feat", _vp), ("dfeat", _vp), ("dz", _vp),
                ("demb_table", _vp * MAX_PLANES), ("donehot", _vp * MAX_PLANES),
                ("dcont", _vp * MAX_PLANES)]


class ActBwdArgs(C.Structure):
    _fields_ = [("d_dtype", _i32), ("d_pitch", _i32), ("y_dtype", _i32), ("y_pitch", _i32),
                ("p_dtype", _i32), ("p_pitch", _i32), ("pixels", _i64), ("pixels_per_sample", _i32),
                ("C", _i32), ("act", _i32), ("slope", _f32), ("mask_pitch", _i32),
                ("bn_mask_pitch", _i32),
                ("dOut", _vp), ("y", _vp), ("dPre", _vp), ("out_mask", _vp), ("dbias", _vp),
                ("bias_mod", _i32),
                ("bn_sums", _vp), ("bn_mask", _vp), ("bn_gamma", _vp), ("bn_mean", _vp),
                ("bn_invstd", _vp), ("bn_dgamma", _vp), ("bn_dbeta", _vp), ("bn_inv_world", _f32)]


class ScmAffineArgs(C.Structure):
    _fields_ = [("n", _i64), ("hidden", _i32), ("closed", _f32 * 3), ("clip_lo", _f32), ("clip_hi", _f32),
                ("lo", _f32), ("span", _f32), ("u_min", _f32), ("u_max", _f32), ("parent_shift", _f32),
                ("v_min", _f32), ("v_max", _f32), ("p_min", _f32), ("p_max", _f32),
                ("w1", _vp), ("b1", _vp), ("w2", _vp), ("b2", _vp),
                ("value", _vp), ("parent", _vp), ("parent_cf", _vp), ("noise_out", _vp), ("value_cf", _vp),
                ("parent_cf_out", _vp), ("value_cf_scaled", _vp), ("parent_cf_scaled", _vp)]


class ExplainGroup(C.Structure):
    _fields_ = [("mode", _i32), ("width", _i32), ("offset", _i64), ("dout", _vp)]


_SIGS = {
    "icf_last_error": (C.c_char_p, []),
    "icf_version": (_i32, []),
    "icf_tc_enabled": (_i32, []),
    "icf_set_tc_enabled": (None, [_i32]),
    "icf_last_conv_path": (_i32, []),
    "icf_ws_plan": (_i32, [C.POINTER(ConvArgs), C.POINTER(C.c_int32), _i32]),
    "icf_conv_forward": (_i32, [C.POINTER(ConvArgs), _vp]),
    "icf_conv_forward_splitk": (_i32, [C.POINTER(ConvArgs), _vp, _i64, _vp]),
    "icf_conv_wgrad": (_i32, [C.POINTER(WgradArgs), _vp]),
    "icf_wgrad_plan": (_i32, [C.POINTER(WgradArgs), C.POINTER(C.c_int32), _i32]),
    "icf_pack": (_i32, [_vp, _vp, _i32, C.POINTER(Perm), _vp]),
    "icf_pack_multi": (_i32, [_vp, _i32, _i64, _vp]),
    "icf_unpack_multi": (_i32, [_vp, _i32, _i64, _vp]),
    "icf_unpack": (_i32, [_vp, _vp, C.POINTER(Perm), _i32, _vp]),
    "icf_pack4": (_i32, [_vp, _vp, _i32, C.POINTER(Perm4), _vp]),
    "icf_unpack4": (_i32, [_vp, _vp, C.POINTER(Perm4), _vp]),
    "icf_argmax_rows": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "icf_image_features_fwd": (_i32, [C.POINTER(ImgFeatArgs), _vp]),
    "icf_image_features_bwd": (_i32, [C.POINTER(ImgFeatArgs), _vp]),
    "icf_latent_features_fwd": (_i32, [C.POINTER(LatFeatArgs), _vp]),
    "icf_latent_features_bwd": (_i32, [C.POINTER(LatFeatArgs), _vp]),
    "icf_bn_finalize": (_i32, [_vp, _i32, C.c_double, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp]),
    "icf_scale_shift_mask": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _vp, _vp, _vp,
                                    _i32, _vp]),
    "icf_bn_bwd_reduce": (_i32, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _vp, _i32, _vp, _vp,
                                 _vp, _vp]),
    "icf_act_backward": (_i32, [C.POINTER(ActBwdArgs), _vp]),
    "icf_bce_logits": (_i32, [_vp, _i32, _i32, _i32, _f32, _f32, _vp, _vp, _i32, _i32, _vp]),
    "icf_sigmoid_mean": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "icf_adam_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "icf_mse_loss": (_i32, [_vp, _i64, _vp, _i32, _i32, _i64, _i64, _f32, _vp, _vp, _vp, _i32, _i32, _vp]),
    "icf_col_mean": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp]),
    "icf_latent_l2": (_i32, [_vp, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _i32, _vp]),
    "icf_scm_affine_cf": (_i32, [C.POINTER(ScmAffineArgs), _vp]),
    "icf_onehot_swap": (_i32, [_vp, _i32, _vp, _i64, _i32, _vp, _vp]),
    "icf_explain_transform": (_i32, [_vp, _vp, C.POINTER(ExplainGroup), _i32, _i64, _vp]),
    "icf_explain_backward": (_i32, [_vp, _vp, C.POINTER(ExplainGroup), _i32, _i64, _vp]),
    "icf_log_spectrogram": (_i32, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _i32, _vp]),
    "icf_spect_stats": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "icf_spect_to_img": (_i32, [_vp, _vp, _vp, _i64, _i32, _f32, _vp, _i32, _vp]),
    "icf_col2im_taps": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _f32, _vp,
                                _i32, _i32, _vp]),
    "icf_im2col_taps": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "icf_bn_fold_weights": (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "icf_bn_fold_wgrad": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "icf_workspace_bytes": (_i64, [C.c_char_p, _vp]),
    "icf_cast": (_i32, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "icf_fill_f32": (_i32, [_vp, _f32, _i64, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)

_lib = None


def _check_fresh():
    """Refuse a library that was not built from the sources of this tree (build.py stores their fingerprint next to it)."""
    build_py = os.path.join(os.path.dirname(_HERE), "build.py")
    if os.environ.get("ICF_SKIP_FRESH_CHECK") or not os.path.exists(build_py):
        return
    import importlib.util
    spec = importlib.util.spec_from_file_location("icf_build", build_py)
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    if b.built_hash() != b.source_hash():
        raise RuntimeError(
            f"{LIB_PATH} is stale: it was not built from the current csrc/ + include/icf.h "
            "(run `python imagecfgen-pytorch_b200/build.py` or __graft_entry__.build()).")


def load():
    """Load (once) and return the ctypes handle; raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python imagecfgen-pytorch_b200/build.py` "
            "(or __graft_entry__.build()). The hot path has no CPU / eager fallback.")
    _check_fresh()
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what=""):
    if status != 0:
        msg = load().icf_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libicf_b200 {what} failed (status {status}): {msg}")
