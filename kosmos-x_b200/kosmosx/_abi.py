"""ctypes binding of libkosmosx_sm100.so (the C ABI declared in include/kosmosx_b200.h).

There is no fallback: if the shared library is missing, importing this module raises, and
every entry point returns an error (surfaced as RuntimeError) when no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# KX_LIB selects another build of the same library (the -DKX_GEMM_TRACE profiling variant); default = the shipped one
LIB_PATH = os.environ.get("KX_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libkosmosx_sm100.so")

KX_OK = 0
KX_ACT_NONE, KX_ACT_GELU, KX_ACT_QUICK_GELU = 0, 1, 2
KX_EPI_GENERIC, KX_EPI_QKV_XPOS = 0, 1
KX_MAX_IMAGES = 16
KX_DEC_PLAIN, KX_DEC_RESIDUAL, KX_DEC_QKV = 0, 1, 2
KX_DECODE_MAX_BATCH = 32
KX_LOSS_REFERENCE, KX_LOSS_NEXT_TOKEN = 0, 1
KX_SUMSQ_SCRATCH = 2048

_f32p = C.c_void_p
_vp = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_f = C.c_float


class GemmArgs(C.Structure):
    """struct kx_gemm_args (include/kosmosx_b200.h)."""

    _fields_ = [
        ("M", _i), ("N", _i), ("K", _i),
        ("bias", _f32p), ("res", _f32p), ("ld_res", _ll),
        ("out", _vp), ("ld_out", _ll), ("out_f32", _i), ("act", _i), ("epi", _i),
        ("grp_rows", _i), ("grp_stride", _i), ("grp_off", _i),
        ("add_tab", _f32p), ("add_off", _i), ("ld_add", _ll),
        ("xq_cos", _f32p), ("xq_sin", _f32p), ("xk_cos", _f32p), ("xk_sin", _f32p),
        ("seq_len", _i), ("d_model", _i),
        ("cta_group", _i), ("block_n", _i), ("max_ctas", _i), ("epi_mode", _i),
        ("ln_part", _f32p), ("ln_c", _f32p), ("ln_tiles", _i), ("ln_cols", _i), ("ln_eps", _f),
        ("stats_out", _f32p), ("out2", _vp), ("ld_out2", _ll),
        ("a_trans", _i), ("b_trans", _i),
        ("drop_p", _f), ("drop_site", C.c_uint), ("drop_seed", C.c_ulonglong),
    ]


class DecodeLinearArgs(C.Structure):
    """struct kx_decode_linear_args (include/kosmosx_b200.h)."""

    _fields_ = [
        ("mode", _i), ("act", _i), ("bias", _f32p), ("ln_c", _f32p), ("ln_eps", _f),
        ("out", _vp), ("ld_out", _ll), ("out_f32", _i), ("argmax_keys", _vp),
        ("x", _f32p), ("ld_x", _ll), ("xb", _vp), ("ld_xb", _ll),
        ("q_out", _vp), ("ld_q", _ll), ("k_cache", _vp), ("v_cache", _vp),
        ("t_max", _i), ("d_model", _i), ("pos", _vp),
        ("xq_cos", _f32p), ("xq_sin", _f32p), ("xk_cos", _f32p), ("xk_sin", _f32p),
    ]


_pp = C.POINTER(C.c_void_p)


class DecodeStepArgs(C.Structure):
    """struct kx_decode_step_args (include/kosmosx_b200.h)."""

    _fields_ = [
        ("batch", _i), ("layers", _i), ("d_model", _i), ("ffn", _i), ("heads", _i), ("vocab", _i), ("t_max", _i), ("pos_rows", _i),
        ("text_index_off", _i),
        ("eps", _f), ("scale", _f),
        ("w_qkv", _pp), ("c_qkv", _pp), ("d_qkv", _pp),
        ("w_o", _pp), ("c_o", _pp), ("d_o", _pp),
        ("w_fc1", _pp), ("c_fc1", _pp), ("d_fc1", _pp),
        ("w_fc2", _pp), ("c_fc2", _pp), ("d_fc2", _pp),
        ("k_cache", _pp), ("v_cache", _pp),
        ("w_out", _vp), ("c_out", _f32p), ("d_out", _f32p),
        ("embed_table", _f32p), ("pos_table", _f32p),
        ("xq_cos", _f32p), ("xq_sin", _f32p), ("xk_cos", _f32p), ("xk_sin", _f32p),
        ("tokens", _vp), ("x", _f32p), ("xb", _vp), ("q", _vp), ("att", _vp), ("mid", _vp),
        ("logits", _f32p), ("ld_logits", _ll),
        ("argmax_keys", _vp), ("pos", _vp), ("step", _vp), ("err_flag", _vp),
        ("forced", _vp), ("history", _vp), ("history_ld", _i),
        ("barrier", _vp), ("trace", _vp),
    ]


# name -> (restype, argtypes); must list every symbol of include/kosmosx_b200.h
SIGNATURES = {
    "kx_last_error": (C.c_char_p, []),
    "kx_abi_version": (_i, []),
    "kx_device_check": (_i, []),
    "kx_launch_count": (C.c_ulonglong, []),
    "kx_gemm_bf16": (_i, [_vp, _ll, _vp, _ll, C.POINTER(GemmArgs), _vp]),
    "kx_attn_fwd": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _f32p, _vp]),
    "kx_attn_set_trace": (_i, [_vp]),
    "kx_rowstats_cast": (_i, [_f32p, _ll, _vp, _ll, _f32p, _i, _i, _vp]),
    "kx_perceiver_xattn_fwd": (_i, [_vp, _ll, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _f, _vp]),
    "kx_layernorm_fwd": (_i, [_vp, _i, _ll, _f32p, _i, _i, _f32p, _f32p, _f, _vp, _i, _ll, _i, _i, _i, _i, _i, _vp]),
    "kx_add_positions": (_i, [_f32p, _f32p, _i, _i, _i, _f32p, _i, _vp]),
    "kx_embed_splice_pos": (_i, [_vp, _i, _i, _f32p, _i, _f32p, _i, _i, C.POINTER(_i), _i, _i, _i, _f32p, _vp, _vp]),
    "kx_im2col_patches": (_i, [_f32p, _i, _i, _i, _i, _vp, _i, _f32p, _f32p, _f32p, _i, _vp]),
    "kx_clip_normalize_u8": (_i, [_vp, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _f32p, _vp]),
    "kx_im2col_patches_u8": (_i, [_vp, _i, C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _i, _vp, _i, _f32p, _f32p, _f32p, _i, _vp]),
    "kx_resize_crop_u8": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kx_xpos_tables": (_i, [_f32p, _f32p, _i, _i, _f, _f32p, _f32p, _f32p, _f32p, _vp]),
    "kx_cast_f32_to_bf16": (_i, [_f32p, _vp, _ll, _vp]),
    "kx_cast_bf16_to_f32": (_i, [_vp, _f32p, _ll, _vp]),
    "kx_broadcast_rows": (_i, [_f32p, _f32p, _ll, _i, _vp]),
    # ---- verification precision (bf16x3)
    "kx_split_bf16x3": (_i, [_f32p, _ll, _i, _i, _i, _vp, _ll, _i, _vp]),
    "kx_attn_f32": (_i, [_f32p, _ll, _f32p, _f32p, _ll, _f32p, _ll, _i, _i, _i, _i, _i, _f, _vp]),
    "kx_xpos_apply_f32": (_i, [_f32p, _ll, _i, _i, _i, _f32p, _f32p, _f32p, _f32p, _vp]),
    "kx_im2col_patches_f32": (_i, [_f32p, _i, _i, _i, _i, _f32p, _i, _f32p, _f32p, _f32p, _i, _vp]),
    # ---- training step
    "kx_attn_fwd_lse": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _f32p, _f32p, _vp]),
    "kx_attn_bwd": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _f32p, _vp, _vp, _vp, _ll, _f32p, _f32p,
                         _f32p, _f32p, _f32p, _f32p, _i, _i, _i, _i, _f, _vp]),
    "kx_attn_bwd_set_trace": (_i, [_vp]),
    "kx_act_layernorm_fwd": (_i, [_vp, _ll, _i, _f32p, _f32p, _f, _vp, _ll, _i, _i, _vp]),
    "kx_ln_bwd_partials": (_i, [_i]),
    "kx_layernorm_bwd": (_i, [_vp, _i, _ll, _f32p, _i, _vp, _ll, _f32p, _f, _f32p, _ll, _vp, _i, _ll, _vp, _ll, _f32p, _i,
                              _f32p, _f32p, _f32p, _i, _i, _i, _f, C.c_uint, C.c_ulonglong, _vp]),
    "kx_dropout_f32": (_i, [_f32p, _ll, _i, _i, _f, C.c_uint, C.c_ulonglong, _vp]),
    "kx_attn_dropout_mask_words": (C.c_size_t, [_i, _i, _i]),
    "kx_attn_dropout_masks": (_i, [_f, C.c_uint, C.c_ulonglong, _i, _i, _i, _i, _vp, _vp, _vp]),
    "kx_attn_fwd_dropout": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _f32p, _f32p, _f, _vp, _vp]),
    "kx_attn_bwd_dropout": (_i, [_vp, _vp, _vp, _ll, _vp, _ll, _vp, _ll, _f32p, _vp, _vp, _vp, _ll, _f32p, _f32p,
                                 _f32p, _f32p, _f32p, _f32p, _i, _i, _i, _i, _f, _f, _vp, _vp]),
    "kx_perceiver_xattn_bwd": (_i, [_vp, _ll, _vp, _ll, _i, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _i, _i, _f, _vp]),
    "kx_gelu_fwd": (_i, [_vp, _vp, _ll, _vp]),
    "kx_gelu_bwd": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "kx_act_fwd": (_i, [_vp, _vp, _ll, _i, _vp]),
    "kx_act_bwd": (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    "kx_gather_rows": (_i, [_vp, _i, _ll, _vp, _ll, _i, _i, _i, _i, _i, _i, _vp]),
    "kx_sum_rows_f32": (_i, [_f32p, _ll, _i, _ll, _f32p, _i, _vp]),
    "kx_colsum_bf16": (_i, [_vp, _ll, _i, _i, _f32p, _vp]),
    "kx_xpos_bwd": (_i, [_vp, _ll, _i, _i, _i, _f32p, _f32p, _f32p, _f32p, _vp]),
    "kx_loss_targets": (_i, [_vp, _i, _i, C.POINTER(_i), _i, _i, _i, _ll, _vp, _f32p, _vp]),
    "kx_ce_fwd_bwd": (_i, [_f32p, _ll, _vp, _i, _i, _f32p, _vp, _ll, _f32p, _vp, _vp]),
    "kx_embed_bwd": (_i, [_f32p, _vp, _i, _i, C.POINTER(_i), _i, _i, _i, _i, _i, _i, _f32p, _f32p, _vp]),
    "kx_sumsq": (_i, [_f32p, _ll, _f32p, _f32p, _vp]),
    "kx_clip_scale": (_i, [_f32p, _f, _f, _f32p, _f32p, _vp]),
    "kx_adamw_step": (_i, [_f32p, _f32p, _f32p, _f32p, _vp, _ll, _f, _f, _f, _f, _f, _i, _f32p, _vp]),
    "kx_lion_step": (_i, [_f32p, _f32p, _f32p, _vp, _ll, _f, _f, _f, _f, _f32p, _vp]),
    # ---- incremental decoding
    "kx_decode_linear": (_i, [_vp, _ll, _i, _vp, _ll, _i, _i, C.POINTER(DecodeLinearArgs), _vp]),
    "kx_decode_attn_scratch_bytes": (C.c_size_t, [_i, _i, _i]),
    "kx_decode_attn": (_i, [_vp, _ll, _vp, _vp, _i, _i, _i, _vp, _f, _f32p, _vp, _vp, _ll, _vp]),
    "kx_kv_cache_store": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _i, _vp]),
    "kx_decode_embed": (_i, [_vp, _i, _f32p, _i, _f32p, _i, _vp, _i, _i, _f32p, _vp, _vp, _vp]),
    "kx_argmax_advance": (_i, [_f32p, _ll, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "kx_decode_plan_bytes": (C.c_size_t, [_i]),
    "kx_decode_step_ctas": (_i, []),
    "kx_decode_plan_build": (_i, [C.POINTER(DecodeStepArgs), _vp, _vp]),
    "kx_decode_step": (_i, [_vp, _vp]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with kosmos-x_b200/build.sh (or __graft_entry__.build()). "
            "kosmosx has no CPU or PyTorch fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library mismatch
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def last_error() -> str:
    return (lib.kx_last_error() or b"").decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    if status != KX_OK:
        raise RuntimeError(f"{what} failed (status {status}): {last_error()}")
