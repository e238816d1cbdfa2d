"""ctypes binding of libivgpt_b200.so (C ABI declared in include/ivgpt_b200.h).

The product path has NO fallback: if the shared library is missing, or a tensor is not on a CUDA device,
the call raises.  Importing this module never touches the GPU.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IVGPT_B200_LIB: another build of the same library (A/B runs of compile-time variants); default = the in-tree build
LIB_PATH = os.environ.get("IVGPT_B200_LIB") or os.path.join(_HERE, "libivgpt_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_SWIGLU = 0, 1, 2


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("bn", C.c_int),
        ("a", C.c_void_p), ("lda", C.c_longlong), ("a_bstride", C.c_longlong),
        ("a_rows", C.c_int), ("a_cols", C.c_int), ("a_batches", C.c_int),
        ("b", C.c_void_p), ("ldb", C.c_longlong), ("b_bstride", C.c_longlong),
        ("b_rows", C.c_int), ("b_cols", C.c_int), ("b_batches", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("batch", C.c_int), ("heads", C.c_int),
        ("a_bsel", C.c_int), ("a_bdiv", C.c_int), ("b_bsel", C.c_int), ("b_bdiv", C.c_int), ("o_bsel", C.c_int),
        ("a_kbase", C.c_int), ("a_khead", C.c_int), ("b_kbase", C.c_int), ("b_khead", C.c_int),
        ("b_nhead", C.c_int), ("o_nhead", C.c_int),
        ("causal_skip", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_longlong), ("out_bstride", C.c_longlong), ("out_dtype", C.c_int),
        ("bias", C.c_void_p), ("bias_along_m", C.c_int),
        ("residual", C.c_void_p), ("ldr", C.c_longlong), ("res_bstride", C.c_longlong), ("res_dtype", C.c_int),
        ("act", C.c_int),
        ("alpha", C.c_float),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("dtype", C.c_int), ("bn", C.c_int),
        ("x", C.c_void_p), ("N", C.c_int), ("Hin", C.c_int), ("Win", C.c_int), ("Cin", C.c_int), ("stride", C.c_int),
        ("w", C.c_void_p), ("Cout", C.c_int),
        ("x2", C.c_void_p), ("C2", C.c_int),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("res_dtype", C.c_int),
        ("act", C.c_int),
        ("out", C.c_void_p), ("out_dtype", C.c_int),
        ("gn_part", C.c_void_p), ("gn_groups", C.c_int),
        ("in_scale", C.c_void_p), ("in_shift", C.c_void_p), ("in_silu", C.c_int),
    ]


class MegaDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("hidden", C.c_int), ("inter", C.c_int), ("heads", C.c_int), ("layers", C.c_int),
        ("vocab", C.c_int), ("Lmax", C.c_int), ("steps", C.c_int),
        ("o_splits", C.c_int), ("d_splits", C.c_int),
        ("eps", C.c_float),
        ("x", C.c_void_p), ("xn", C.c_void_p), ("qkv", C.c_void_p), ("ao", C.c_void_p), ("act", C.c_void_p),
        ("part", C.c_void_p), ("logits", C.c_void_p),
        ("ldl", C.c_longlong),
        ("kcache", C.c_void_p), ("vcache", C.c_void_p),
        ("embed", C.c_void_p), ("norm_f", C.c_void_p), ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
        ("tokens", C.c_void_p), ("tok_stride", C.c_longlong),
        ("dpos", C.c_void_p),
        ("do_sample", C.c_int), ("topk", C.c_int), ("inv_temp", C.c_float),
        ("dseed", C.c_void_p),
        ("barrier", C.c_void_p), ("error", C.c_void_p),
        ("layers_dev", C.c_void_p), ("lm_head_packed", C.c_void_p),
        ("prof", C.c_void_p),
        ("vrows", C.c_void_p),
        ("attn_part", C.c_void_p), ("attn_cnt", C.c_void_p),
        ("attn_mode", C.c_int),
        ("slot0", C.c_int), ("slot_period", C.c_int), ("nslots", C.c_int),
        ("slot_token", C.c_longlong),
        ("slot_emb", C.c_void_p),
        ("a_bulk", C.c_int),
        ("mma_m64", C.c_int),
        ("gemm_mode", C.c_int), ("qkv_splits", C.c_int), ("a_rows", C.c_int),
        ("qkvp", C.c_void_p),
        ("bn_wide", C.c_int),
        ("bn_down", C.c_int),
        ("tile_cnt", C.c_void_p),
    ]


# name -> argtypes (return type is always int unless listed in _RESTYPES)
_P, _I, _L, _F, _U = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong
SIGNATURES = {
    "ivgpt_last_error": [],
    "ivgpt_launch_count": [],
    "ivgpt_device_info": [C.POINTER(C.c_int)] * 3,
    "ivgpt_profile_enable": [_I],
    "ivgpt_profile_collect": [_I, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong)],
    "ivgpt_count_add": [_L],
    "ivgpt_vq_argmin": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ivgpt_gemm": [C.POINTER(GemmDesc), _P],
    "ivgpt_conv3x3": [C.POINTER(ConvDesc), _P],
    "ivgpt_conv3x3_plan": [C.POINTER(ConvDesc), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "ivgpt_groupnorm_finalize": [_P, _P, _I, _I, _I, C.c_double, _F, _P],
    "ivgpt_groupnorm_coeff": [_P, _P, _P, _P, _P, _I, _I, _I, _P],
    "ivgpt_groupnorm_stats": [_I, _P, _P, _P, _I, _I, _I, _I, _F, _P],
    "ivgpt_groupnorm_apply": [_I, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _I, _I, _P],
    "ivgpt_conv_in": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P],
    "ivgpt_conv_out3": [_I, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P],
    "ivgpt_upsample2x": [_I, _P, _P, _I, _I, _I, _I, _P],
    "ivgpt_patchify": [_I, _P, _P, _I, _I, _I, _I, _I, _P],
    "ivgpt_convert": [_I, _P, _I, _P, _L, _P],
    "ivgpt_vq_commit": [_I, _P, _P, _P, _P, _L, _I, _L, _F, _P, _P, _P],
    "ivgpt_tokens_serialise": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _L, _L, _P],
    "ivgpt_tokens_gather": [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _L, _L, _I, _P, _P],
    "ivgpt_embed": [_P, _L, _I, _P, _P, _P, _L, _I, _L, _P],
    "ivgpt_add_rows": [_P, _P, _L, _P],
    "ivgpt_rmsnorm": [_I, _P, _P, _P, _L, _I, _F, _P],
    "ivgpt_rope_kv": [_I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "ivgpt_softmax": [_I, _P, _P, _L, _I, _I, _L, _L, _I, _I, _P],
    "ivgpt_decode_attn": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _F, _P],
    "ivgpt_argmax": [_P, _L, _I, _I, _P, _L, _P, _P],
    "ivgpt_topk_sample": [_P, _L, _I, _I, _I, _F, _U, _U, _P, _L, _P, _P, _P],
    "ivgpt_ce_loss": [_P, _L, _I, _I, _I, _P, _P, _P, _P, _P],
    "ivgpt_incr": [_P, _I, _P],
    "ivgpt_preprocess_resize": [_I, _P, _L, _L, _L, _L, _I, _I, _I, _I, _P, _I, _I, _F, _P],
    "ivgpt_slot_embed_add": [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    "ivgpt_slot_force": [_P, _L, _P, _I, _I, _I, _L, _P],
    "ivgpt_decode_attn_fused": [_I, _P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _F, _P],
    "ivgpt_set_pdl": [_I],
    "ivgpt_transpose": [_I, _P, _P, _I, _I, _I, _L, _L, _L, _L, _P],
    "ivgpt_swiglu": [_I, _I, _P, _P, _P, _L, _P],
    "ivgpt_rmsnorm_bwd": [_I, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P],
    "ivgpt_softmax_bwd": [_I, _P, _P, _P, _L, _I, _I, _L, _I, _F, _P],
    "ivgpt_rope_bwd": [_I, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P],
    "ivgpt_ce_bwd": [_I, _P, _L, _I, _I, _I, _P, _P, _F, _P, _L, _P],
    "ivgpt_embed_bwd": [_P, _P, _P, _L, _I, _L, _P],
    "ivgpt_adamw": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P],
    "ivgpt_add_to_f32": [_I, _P, _P, _L, _P],
    "ivgpt_dropout": [_I, _P, _P, _L, _F, _U, _P],
    "ivgpt_colsum": [_P, _L, _I, _L, _P, _I, _P, _I, _P],
    "ivgpt_groupnorm_bwd_chunks": [_I, _I],
    "ivgpt_groupnorm_bwd": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P],
    "ivgpt_im2col3x3_t": [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    "ivgpt_transpose_pad": [_P, _P, _I, _I, _I, _I, _I, _I, _L, _L, _P],
    "ivgpt_zero_insert2x": [_P, _P, _I, _I, _I, _I, _P],
    "ivgpt_upsample2x_bwd": [_P, _P, _I, _I, _I, _I, _P],
    "ivgpt_silu": [_P, _P, _P, _L, _P],
    "ivgpt_axpby": [_P, _P, _P, _F, _F, _L, _P],
    "ivgpt_reduce_mid": [_P, _P, _L, _I, _L, _I, _P],
    "ivgpt_nchw_to_nhwc": [_P, _P, _L, _I, _I, _L, _P],
    "ivgpt_vq_bwd": [_P, _P, _P, _P, _F, _L, _P, _P, _P],
    "ivgpt_vq_set_order": [_I],
    "ivgpt_vq_get_order": [],
    "ivgpt_mega_layer_bytes": [],
    "ivgpt_mega_fill_layer": [_P, _P, _P, _P, _P, _P, _P],
    "ivgpt_mega_packed_elems": [_I, _I],
    "ivgpt_mega_pack_weight": [_P, _P, _I, _I, _P],
    "ivgpt_mega_packed_elems_bn": [_I, _I, _I],
    "ivgpt_mega_pack_weight_bn": [_P, _P, _I, _I, _I, _P],
    "ivgpt_mega_packed_elems64": [_I, _I],
    "ivgpt_mega_pack_weight64": [_P, _P, _I, _I, _I, _P],
    "ivgpt_decode_mega": [C.POINTER(MegaDesc), _P],
    "ivgpt_mega_fused_norm": [],
    "ivgpt_set_deterministic": [_I],
    "ivgpt_set_mma_issuers": [_I],
    "ivgpt_set_gemm_mh2": [_I],
    "ivgpt_flash_attn": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _L, _L, _L, _L, _L, _I, _F, _P],
}
_RESTYPES = {"ivgpt_last_error": C.c_char_p, "ivgpt_launch_count": C.c_ulonglong,
             "ivgpt_mega_packed_elems": C.c_longlong, "ivgpt_mega_packed_elems64": C.c_longlong,
             "ivgpt_mega_packed_elems_bn": C.c_longlong}

_lib = None


class B200LibraryError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or ivideogpt_b200/csrc/build.sh).  ivideogpt_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def check(status: int, what: str):
    if status != 0:
        msg = load().ivgpt_last_error()
        raise B200LibraryError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def launch_count() -> int:
    return int(load().ivgpt_launch_count())
