"""ctypes binding of libcorrif_b200.so (the C ABI declared in include/corrif.h).

There is no CPU fallback: if the library is missing or the device is not sm_100, every entry point
raises.  ``load()`` only dlopens the library (safe without a GPU); compute needs a B200.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libcorrif_b200.so")

f32p = C.c_void_p      # device pointers travel as integers
u8p = C.c_void_p
f64p = C.c_void_p
u64p = C.c_void_p
i32, i64, u32, u64, f32 = C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float
stream_t = C.c_void_p


class GemmDesc(C.Structure):
    """Mirror of ``corrif_gemm_desc`` (include/corrif.h)."""
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("aux", C.c_void_p),
        ("lda", i64), ("ldb", i64), ("ldd", i64), ("ldr", i64), ("ldaux", i64),
        ("M", i32), ("N", i32), ("K", i32),
        ("a_mn_major", i32), ("b_mn_major", i32),
        ("batch_outer", i32), ("batch_inner", i32),
        ("a_bo", i64), ("a_bi", i64), ("b_bo", i64), ("b_bi", i64), ("d_bo", i64), ("d_bi", i64),
        ("split_k", i32), ("epilogue", i32), ("precision", i32), ("alpha", f32),
        ("flags", i32),
        ("drop_p", f32), ("drop_site_a", u32), ("drop_site_b", u32), ("drop_site_bo", u32),
        ("drop_seed", u64), ("drop_seed_dev", C.c_void_p), ("bias_bo", i64),
    ]


class VolSrc(C.Structure):
    """Mirror of ``corrif_vol_src``."""
    _fields_ = [("p", C.c_void_p), ("C", i32), ("reserved", i32), ("ld", i64)]


class Conv3dDesc(C.Structure):
    """Mirror of ``corrif_conv3d_desc`` (include/corrif.h)."""
    _fields_ = [
        ("src", VolSrc * 3), ("nsrc", i32), ("B", i32), ("D", i32), ("H", i32), ("W", i32),
        ("Cin", i32), ("Cout", i32), ("ksize", i32), ("pad_mode", i32), ("relu", i32),
        ("wpk", C.c_void_p), ("bias", C.c_void_p), ("out", C.c_void_p), ("ldo", i64), ("stats", C.c_void_p),
    ]


PAD_ZEROS, PAD_REPLICATE, PAD_REPLICATE_ADJOINT = 0, 1, 2
EPI_STORE, EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_MUL_DGELU, EPI_ATOMIC_ADD = range(6)
GEMM_TF32, GEMM_FP32 = 0, 1
GEMM_ROUND_TF32 = 1
NO_SITE = 0xFFFFFFFF

# name -> (restype, argtypes); must list every symbol of include/corrif.h (tests check this)
PROTOTYPES = {
    "corrif_abi_version": (C.c_int, []),
    "corrif_last_error": (C.c_char_p, []),
    "corrif_check_device": (C.c_int, []),
    "corrif_gemm": (C.c_int, [C.POINTER(GemmDesc), stream_t]),
    "corrif_sizeof_gemm_desc": (C.c_int, []),
    "corrif_transpose": (C.c_int, [f32p, f32p, i64, i32, i32, i32, stream_t]),
    "corrif_round_tf32": (C.c_int, [f32p, f32p, i64, stream_t]),
    "corrif_round_tf32_multi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, i32, stream_t]),
    "corrif_layernorm_fwd": (C.c_int, [f32p, f32p, i64, f32p, f32p, f32p, f32p, f32p, f32p, i64, i32, i32, stream_t]),
    "corrif_layernorm_bwd_scratch_floats": (i64, [i64, i32]),
    "corrif_layernorm_bwd": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, i64, i32, i32,
                                       f32p, f32, u64, u64p, u32, u32, stream_t]),
    "corrif_layernorm_bwd_regroup": (C.c_int, [f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, f32p, i64, i32, i32,
                                               f32p, f32, u64, u64p, u32, u32, i32, i32, f32p, stream_t]),
    "corrif_softmax_fwd": (C.c_int, [f32p, f32p, i64, i32, f32, u64, u64p, u32, i32, stream_t]),
    "corrif_softmax_bwd": (C.c_int, [f32p, f32p, i64, i32, f32, f32, u64, u64p, u32, stream_t]),
    "corrif_attention_fwd": (C.c_int, [f32p, f32p, f32p, C.c_void_p, i32, i32, i32, i32, f32, f32, u64, u64p, u32, i32, u32, i32, stream_t]),
    "corrif_attention_keepbits": (C.c_int, [C.c_void_p, i32, i32, i32, f32, u64, u64p, u32, i32, u32, i32, stream_t]),
    "corrif_attention_fwd_premasked": (C.c_int, [f32p, f32p, f32p, C.c_void_p, i32, i32, i32, i32, f32, f32, i32, stream_t]),
    "corrif_attention_bwd": (C.c_int, [f32p, f32p, f32p, f32p, C.c_void_p, f32p, f32p, i32, i32, i32, i32, f32, f32, stream_t]),
    "corrif_dropout": (C.c_int, [f32p, f32p, i64, f32, u64, u64p, u32, stream_t]),
    "corrif_dropout_colsum": (C.c_int, [f32p, f32p, i64, i32, f32, u64, u64p, u32, f32p, stream_t]),
    "corrif_dropout_mask": (C.c_int, [f32p, i64, f32, u64, u64p, u32, stream_t]),
    "corrif_dropout_add": (C.c_int, [f32p, f32p, f32p, i64, f32, u64, u64p, u32, u32, stream_t]),
    "corrif_colsum_scratch_floats": (i64, [i64, i32]),
    "corrif_colsum": (C.c_int, [f32p, i64, i64, i32, f32p, C.c_int, f32p, stream_t]),
    "corrif_colsum_batched": (C.c_int, [f32p, i64, i64, i32, f32p, i32, i64, i64, C.c_int, stream_t]),
    "corrif_batchsum": (C.c_int, [f32p, i64, i64, i64, f32p, C.c_int, stream_t]),
    "corrif_add_rows": (C.c_int, [f32p, i64, f32p, i64, f32p, i64, i64, i32, stream_t]),
    "corrif_inter_corr_fwd": (C.c_int, [f32p, f32p, f32p, i32, i32, i32, i32, stream_t]),
    "corrif_inter_corr_bwd": (C.c_int, [f32p, f32p, f32p, i32, i32, i32, i32, stream_t]),
    "corrif_inter_corr_bwd_layout": (C.c_int, [f32p, f32p, f32p, i32, i32, i32, i32, i32, stream_t]),
    "corrif_jaccard_sums": (C.c_int, [f32p, f32p, i64, f64p, stream_t]),
    "corrif_loss_jaccard_fused": (C.c_int, [f32p, f32p, i64, i32, i64, f32, f64p, f32p, f64p, stream_t]),
    "corrif_jaccard_finish": (C.c_int, [f64p, f32, f32p, stream_t]),
    "corrif_confusion_counts": (C.c_int, [u8p, u8p, i64, i32, u64p, stream_t]),
    "corrif_bce_probs_fwd_bwd": (C.c_int, [f32p, f32p, i64, f32, f64p, f32p, stream_t]),
    "corrif_adam_step": (C.c_int, [f32p, f32p, f32p, f32p, i64, f32, f32, f32, f32, f32, i32, stream_t]),
    "corrif_sizeof_conv3d_desc": (C.c_int, []),
    "corrif_conv3d_pack_floats": (i64, [i32, i32, i32]),
    "corrif_conv3d_pack_weights": (C.c_int, [f32p, f32p, i32, i32, i32, i32, stream_t]),
    "corrif_conv3d_fwd": (C.c_int, [C.POINTER(Conv3dDesc), stream_t]),
    "corrif_conv3d_wgrad": (C.c_int, [C.POINTER(Conv3dDesc), f32p, i64, f32p, stream_t]),
    "corrif_conv3d_tc_supported": (C.c_int, [C.POINTER(Conv3dDesc)]),
    "corrif_conv3d_tc_pack_floats": (i64, [C.POINTER(Conv3dDesc)]),
    "corrif_conv3d_tc_pack_weights": (C.c_int, [C.POINTER(Conv3dDesc), f32p, f32p, i32, stream_t]),
    "corrif_conv3d_tc_fwd": (C.c_int, [C.POINTER(Conv3dDesc), stream_t]),
    "corrif_conv3d_wgrad_tc_supported": (C.c_int, [C.POINTER(Conv3dDesc)]),
    "corrif_conv3d_wgrad_tc": (C.c_int, [C.POINTER(Conv3dDesc), f32p, i64, f32p, stream_t]),
    "corrif_conv3d_dgrad_border": (C.c_int, [f32p, i64, f32p, f32p, i64, i32, i32, i32, i32, i32, i32, stream_t]),
    "corrif_instnorm_apply": (C.c_int, [f32p, i64, f64p, f32p, f32p, i32, i64, i32, f32, stream_t]),
    "corrif_instnorm_bwd_stats": (C.c_int, [f32p, i64, f32p, i64, f64p, i32, i64, i32, stream_t]),
    "corrif_instnorm_relu_bwd_apply": (C.c_int, [f32p, i64, f32p, i64, f32p, f32p, f64p, f32p, i64, f32p, i32, i64, i32,
                                                 i32, stream_t]),
    "corrif_volume_colsum": (C.c_int, [f32p, i64, f32p, i64, i32, stream_t]),
    "corrif_batchnorm_fwd": (C.c_int, [f32p, i64, f64p, f32p, f32p, f32p, i64, f32p, i64, f32p, f32p, f32p, i64, i32, f32, i32,
                                       stream_t]),
    "corrif_batchnorm_bwd_stats": (C.c_int, [f32p, i64, f32p, i64, f32p, i64, f32p, f32p, f64p, i64, i32, i32, stream_t]),
    "corrif_batchnorm_bwd_apply": (C.c_int, [f32p, i64, f32p, i64, f32p, i64, f32p, f32p, f32p, f64p, f32p, i64, f32p, i64,
                                             f32p, f32p, i64, i32, i32, stream_t]),
    "corrif_resize_trilinear_fwd": (C.c_int, [f32p, i64, f32p, i64, i32, i32, i32, i32, i32, i32, i32, i32, stream_t]),
    "corrif_resize_trilinear_bwd": (C.c_int, [f32p, i64, f32p, i64, i32, i32, i32, i32, i32, i32, i32, i32, stream_t]),
    "corrif_conv1_small_fwd": (C.c_int, [f32p, i64, f32p, f32p, f32p, i64, f64p, i32, i64, i32, i32, i32, stream_t]),
    "corrif_conv1_small_wgrad": (C.c_int, [f32p, i64, f32p, i64, f32p, i64, i32, stream_t]),
    "corrif_resize_linear_axis_fwd": (C.c_int, [f32p, f32p, i64, i32, i32, i64, stream_t]),
    "corrif_resize_linear_axis_bwd": (C.c_int, [f32p, f32p, i64, i32, i32, i64, stream_t]),
    "corrif_resize_nearest_fwd": (C.c_int, [f32p, i64, f32p, i64, i32, i32, i32, i32, i32, i32, i32, i32, stream_t]),
    "corrif_resize_nearest_bwd": (C.c_int, [f32p, i64, f32p, i64, i32, i32, i32, i32, i32, i32, i32, i32, stream_t]),
}

_lib = None


class CorrifError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CorrifError(
            "libcorrif_b200.so is not built (%s). Run `python __graft_entry__.py build`; there is no "
            "CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.corrif_abi_version() != 3:
        raise CorrifError("libcorrif_b200.so ABI version mismatch")
    if lib.corrif_sizeof_gemm_desc() != C.sizeof(GemmDesc):
        raise CorrifError("corrif_gemm_desc layout mismatch between include/corrif.h and _lib.py")
    if lib.corrif_sizeof_conv3d_desc() != C.sizeof(Conv3dDesc):
        raise CorrifError("corrif_conv3d_desc layout mismatch between include/corrif.h and _lib.py")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    """Turn a non-zero C-ABI status into a RuntimeError carrying corrif_last_error()."""
    if rc != 0:
        msg = load().corrif_last_error().decode(errors="replace")
        raise CorrifError("%s failed (status %d): %s" % (what or "corrif call", rc, msg))
