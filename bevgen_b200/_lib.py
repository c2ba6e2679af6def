"""ctypes binding of libbevgen_b200.so (the C ABI in include/bevgen_b200.h).

There is deliberately no fallback: if the shared library is missing, or a call returns a negative status, a
RuntimeError is raised.  `load()` works without a GPU (symbol checks); the first compute call needs a B200.
"""
import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["BEVGEN_B200_LIB"]) if os.environ.get("BEVGEN_B200_LIB") else _HERE / "libbevgen_b200.so"      # override: A/B runs of two builds in one gpurun call
MAX_TAPS = 9

GF_GELU, GF_OUT_NCHW, GF_B_MN, GF_CAUSAL_SKIP, GF_CAUSAL_KLIMIT, GF_OUT_T, GF_OUT_F16F8 = 1, 2, 4, 8, 16, 32, 64
PREP_IDENT, PREP_UP2, PREP_S2D = 0, 1, 2


class GemmArgs(C.Structure):
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p),
        ("a_n", C.c_int), ("a_h", C.c_int), ("a_w", C.c_int), ("a_c", C.c_int),
        ("b_hi", C.c_void_p), ("b_lo", C.c_void_p),
        ("b_rows", C.c_int), ("b_cols", C.c_int),
        ("ntaps", C.c_int),
        ("tap_dx", C.c_int * MAX_TAPS), ("tap_dy", C.c_int * MAX_TAPS), ("tap_dn", C.c_int * MAX_TAPS),
        ("a_n_mul", C.c_int), ("a_n_zstride", C.c_int),
        ("k", C.c_int),
        ("a_c_off", C.c_int), ("a_c_zstride", C.c_int),
        ("b_k_off", C.c_int), ("b_k_zstride", C.c_int),
        ("b_row_zstride", C.c_int), ("b_row_tapstride", C.c_int),
        ("z_inner", C.c_int), ("z_outer", C.c_int),
        ("tile_w", C.c_int), ("tile_h", C.c_int),
        ("out_w", C.c_int), ("out_h", C.c_int), ("n_cols", C.c_int),
        ("out_zo_stride", C.c_longlong), ("out_zi_stride", C.c_longlong),
        ("ldc", C.c_int),
        ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("out_f32", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("flags", C.c_int),
        ("causal_ncond", C.c_int),
        ("bn", C.c_int),
        ("npass", C.c_int),
        ("fin_mode", C.c_int), ("fin_gelu", C.c_int), ("fin_rows", C.c_int),
        ("fin_bias", C.c_void_p), ("fin_resid", C.c_void_p),
        ("fin_x", C.c_void_p), ("fin_y", C.c_void_p), ("fin_hi", C.c_void_p), ("fin_lo", C.c_void_p),
        ("fin_gamma", C.c_void_p), ("fin_beta", C.c_void_p),
        ("fin_eps", C.c_float),
        ("fin_counters", C.c_void_p),
        ("lo_scale", C.c_float),
    ]


class EmbedArgs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("cam_idx", "bev_idx", "intrinsics_inv", "extrinsics_inv", "x_tok_emb", "cond_tok_emb",
                                          "x_pos_emb", "cond_static", "img_embed_w", "cam_embed_w", "forward_shuffle_idx", "pixel", "out",
                                          "step_ptr")] + \
               [(n, C.c_int) for n in ("B", "ncam", "hw", "nc", "n_img", "L", "d", "vocab", "pad_last", "bev_embed", "row0", "nrows")]


class DecodeLayer(C.Structure):
    """bevgen_decode_layer (include/bevgen_b200.h): one entry per transformer block, uploaded as a device array."""
    _fields_ = [(n, C.c_void_p) for n in ("w_qkv", "w_1", "w_2", "c1_qkv", "c2_qkv", "c2_2", "ln1_g", "ln1_b", "c1_1", "c2_1", "k_cache", "v_cache",
                                          "layout")] + [(n, C.c_float) for n in ("s_qkv", "s_1", "s_2", "pad_")]


class DecodeArgs(C.Structure):
    """bevgen_decode_args (include/bevgen_b200.h)."""
    _fields_ = [("layers", C.c_void_p), ("n_layers", C.c_int), ("w_head", C.c_void_p), ("s_head", C.c_float), ("c1_head", C.c_void_p), ("c2_head", C.c_void_p)] + \
               [(n, C.c_int) for n in ("batch", "d", "heads", "vocab", "n_cond", "n_img", "lmax", "ncam", "hw", "step_begin", "step_end")] + \
               [(n, C.c_void_p) for n in ("cam_idx", "x_tok_emb", "x_pos_emb", "img_embed_w", "cam_embed_w", "intrinsics_inv", "extrinsics_inv", "pixel",
                                          "forward_shuffle_idx", "camera_bias")] + \
               [("bias_ld", C.c_int), ("scale", C.c_float), ("temperature", C.c_float), ("top_k", C.c_int), ("greedy", C.c_int),
                ("seed", C.c_ulonglong), ("forced_tokens", C.c_void_p), ("tokens_out", C.c_void_p), ("logits_trace", C.c_void_p),
                ("layout_block", C.c_int), ("layout_ld", C.c_int), ("workspace", C.c_void_p), ("counters", C.c_void_p),
                ("debug", C.c_void_p), ("profile", C.c_void_p)]


_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> (restype, argtypes); must list every symbol declared in include/bevgen_b200.h
SIGNATURES = {
    "bevgen_init": (_i, [_i]),
    "bevgen_last_error": (C.c_char_p, []),
    "bevgen_version": (_i, []),
    "bevgen_sm_count": (_i, []),
    "bevgen_set_pdl": (_i, [_i]),
    "bevgen_gemm_tc": (_i, [C.POINTER(GemmArgs), _vp]),
    "bevgen_conv3x3_halo": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "bevgen_conv3x3_fused": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "bevgen_conv3x3_fused_f16f8": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _i, _vp]),
    "bevgen_groupnorm_affine": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp]),
    "bevgen_groupnorm_finalize": (_i, [_vp, _i, _i, _i, _f, _vp, _vp]),
    "bevgen_groupnorm_stats": (_i, [_vp, _i, _i, _i, _f, _vp, _vp, _vp]),
    "bevgen_prep_operand": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "bevgen_im2col3x3": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "bevgen_transpose_f32": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "bevgen_softmax_rows": (_i, [_vp, _ll, _i, _f, _vp, _vp, _i, _vp]),
    "bevgen_row_sqnorm": (_i, [_vp, _i, _i, _vp, _vp]),
    "bevgen_vq_nearest": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "bevgen_codebook_gather": (_i, [_vp, _vp, _ll, _i, _i, _vp, _vp]),
    "bevgen_conv_in3": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "bevgen_conv_out3": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "bevgen_to_uint8_hwc": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "bevgen_absmax": (_i, [_vp, _ll, _vp, _vp]),
    "bevgen_pack_split_bf16": (_i, [_vp, _ll, _vp, _vp, _vp]),
    "bevgen_pack_f16f8": (_i, [_vp, _ll, _i, _i, _f, _f, _vp, _vp, _vp]),
    "bevgen_denormalize": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "bevgen_layernorm": (_i, [_vp, _ll, _i, _ll, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "bevgen_layernorm_f16f8": (_i, [_vp, _ll, _i, _ll, _vp, _vp, _f, _vp, _vp, _vp, _i, _vp]),
    "bevgen_linear_f16f8": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _f, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "bevgen_ray_embed_add": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "bevgen_mg_head_planes": (_i, [_vp, _ll, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _ll, _i, _i, _i, _vp]),
    "bevgen_mg_sample": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _f, _ll, _vp]),
    "bevgen_mg_remask": (_i, [_vp, _vp, _f, _vp, _vp, _ll, _i, _i, _ll, _vp]),
    "bevgen_mg_geglu_ln": (_i, [_vp, _ll, _vp, _vp, _vp, _ll, _i, _i, _f, _i, _vp]),
    "bevgen_embed_assemble": (_i, [C.POINTER(EmbedArgs), _vp]),
    "bevgen_attn_softmax": (_i, [_vp, _vp, _vp, _ll, _i, _i, _f, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "bevgen_attn_fused_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp]),
    "bevgen_dec_reduce_ln": (_i, [_vp, _i, _ll, _vp, _vp, _ll, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "bevgen_dec_reduce_act": (_i, [_vp, _i, _ll, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "bevgen_kv_store": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "bevgen_dec_attention": (_i, [_vp, _i, _ll, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f,
                                  _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "bevgen_dec_attention_workspace_floats": (_i, [_i, _i]),
    "bevgen_sample_topk": (_i, [_vp, _i, _ll, _i, _i, _f, _i, _i, C.c_ulonglong, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "bevgen_dec_advance": (_i, [_vp, _vp]),
    "bevgen_decode_persistent": (_i, [C.POINTER(DecodeArgs), _vp]),
    "bevgen_decode_workspace": (_i, [_i, _i, _i, _i, C.POINTER(_ll), C.POINTER(_ll)]),
    "bevgen_pack_decode_linear": (_ll, [_vp, _i, _i, _i, _i, _f, _vp, _vp]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing — build it with `python -m bevgen_b200.build` (or __graft_entry__.build()). "
            "bevgen_b200 has no CPU / eager fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().bevgen_last_error().decode(errors="replace")
        raise RuntimeError(f"bevgen_b200: {what} failed with status {status}: {msg}")


_initialised = False


def init():
    """One-time device check (sm_100 only) + driver entry points. Needs a GPU."""
    global _initialised
    if not _initialised:
        check(load().bevgen_init(-1), "bevgen_init")
        _initialised = True
    return load()
