"""ctypes binding of include/mde_b200.h.  Loading fails loudly: there is no Python or CPU fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmde_b200.so")

MDE_FP16, MDE_BF16 = 0, 1
MDE_INPUT_F32_NCHW, MDE_INPUT_U8_HWC = 0, 1
MDE_DT_F32, MDE_DT_U8, MDE_DT_F16, MDE_DT_BF16 = 0, 1, 2, 3
MDE_HEAD_DPT, MDE_HEAD_ENCODER_TAPS, MDE_HEAD_DPT_EXP_SKY = 0, 1, 2
MDE_OUTPUT_MODEL_GRID, MDE_OUTPUT_SOURCE_GRID = 0, 1
MDE_FLAG_SPLIT_K, MDE_FLAG_NO_PDL, MDE_FLAG_NO_GRAPH, MDE_FLAG_SCALE_F32, MDE_FLAG_NORMALISE_F32 = 1, 2, 4, 8, 16
MDE_ABI_VERSION = 2
PRECISIONS = {"fp16": MDE_FP16, "bf16": MDE_BF16}

# every symbol include/mde_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "mde_last_error", "mde_abi_version",
    "mde_engine_create", "mde_engine_set_weight", "mde_engine_load_weights", "mde_engine_finalize",
    "mde_engine_destroy", "mde_engine_num_io", "mde_engine_io_name", "mde_engine_io_shape",
    "mde_engine_io_dtype", "mde_engine_io_is_input", "mde_engine_workspace_bytes",
    "mde_context_create", "mde_context_destroy", "mde_context_set_tensor_address",
    "mde_context_set_input_shape", "mde_context_enqueue", "mde_context_set_gather", "mde_context_launches_per_enqueue",
    "mde_context_get_buffer", "mde_context_snapshot_block", "mde_context_enqueue_timed", "mde_context_op_info",
    "mde_k_preprocess_u8", "mde_k_preprocess_u8_pad", "mde_k_preprocess_u8_square_pad_cubic", "mde_k_preprocess_u8_cubic_f32", "mde_k_im2col_f32", "mde_k_gemm", "mde_k_gemm_tiled", "mde_k_conv3x3", "mde_k_attention", "mde_k_attention_poly", "mde_k_attention_q3", "mde_k_attention_trace", "mde_k_attention_kv", "mde_k_attention_mma",
    "mde_k_layernorm", "mde_k_bilinear", "mde_k_bilinear_add", "mde_k_assemble_tokens", "mde_k_im2col_s2", "mde_k_upconv_head", "mde_k_resize_depth", "mde_k_merge_patches", "mde_k_peer_signal", "mde_k_peer_wait",
    "mde_k_resize_crops", "mde_k_depth_pro_post", "mde_k_resize_depth_halfpixel", "mde_k_resize_depth_halfpixel_nan", "mde_k_qknorm_rope", "mde_k_peer_signal_counter", "mde_k_peer_wait_counter",
]


class EngineDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("embed_dim", C.c_int32), ("depth", C.c_int32),
        ("num_heads", C.c_int32), ("patch_size", C.c_int32), ("features", C.c_int32),
        ("out_channels", C.c_int32 * 4), ("taps", C.c_int32 * 4),
        ("input_h", C.c_int32), ("input_w", C.c_int32), ("batch", C.c_int32),
        ("precision", C.c_int32), ("input_mode", C.c_int32),
        ("max_src_h", C.c_int32), ("max_src_w", C.c_int32), ("swap_rb", C.c_int32),
        ("norm_mean", C.c_double * 3), ("norm_std", C.c_double * 3),
        ("max_depth", C.c_float), ("device", C.c_int32),
        ("output_mode", C.c_int32), ("head_mode", C.c_int32), ("tap_norm_mask", C.c_int32),
        ("num_registers", C.c_int32), ("flags", C.c_int32), ("attn_poly", C.c_int32),
    ]


class Crop(C.Structure):
    _fields_ = [("level_h", C.c_int32), ("level_w", C.c_int32), ("y0", C.c_int32), ("x0", C.c_int32)]


class Epilogue(C.Structure):
    _fields_ = [
        ("d_bias", C.c_void_p), ("d_gamma", C.c_void_p), ("act", C.c_int32),
        ("d_x", C.c_void_p), ("accumulate_x", C.c_int32),
        ("d_res1", C.c_void_p), ("d_res2", C.c_void_p), ("d_out", C.c_void_p), ("d_out_relu", C.c_void_p),
        ("ld_out", C.c_int32), ("tokens", C.c_int32), ("d_pos", C.c_void_p),
        ("shuffle_s", C.c_int32), ("shuffle_cout", C.c_int32), ("shuffle_h", C.c_int32), ("shuffle_w", C.c_int32),
        ("d_head_w", C.c_void_p), ("head_b", C.c_float), ("head_scale", C.c_float), ("d_head_out", C.c_void_p),
        ("gather_n", C.c_int32), ("gather_col0", C.c_int32), ("gather_ld", C.c_int32), ("d_gather", C.c_void_p * 8),
        ("token_skip", C.c_int32), ("head_act", C.c_int32),
    ]


_lib = None


def load() -> C.CDLL:
    """dlopen the in-tree library and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"[MDET] {LIB_PATH} is missing. Build it with `python -m monocular_depth_estimation_trt_b200.build` "
            "(or __graft_entry__.build()). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    P = C.POINTER
    protos = {
        "mde_last_error": (C.c_char_p, []),
        "mde_abi_version": (C.c_int, []),
        "mde_engine_create": (C.c_int, [P(EngineDesc), P(vp)]),
        "mde_engine_set_weight": (C.c_int, [vp, C.c_char_p, vp, i32, P(i64)]),
        "mde_engine_load_weights": (C.c_int, [vp, C.c_char_p]),
        "mde_engine_finalize": (C.c_int, [vp]),
        "mde_engine_destroy": (None, [vp]),
        "mde_engine_num_io": (C.c_int, [vp]),
        "mde_engine_io_name": (C.c_char_p, [vp, i32]),
        "mde_engine_io_shape": (C.c_int, [vp, i32, P(i32), P(i64)]),
        "mde_engine_io_dtype": (C.c_int, [vp, i32]),
        "mde_engine_io_is_input": (C.c_int, [vp, i32]),
        "mde_engine_workspace_bytes": (i64, [vp]),
        "mde_context_create": (C.c_int, [vp, P(vp)]),
        "mde_context_destroy": (None, [vp]),
        "mde_context_set_tensor_address": (C.c_int, [vp, C.c_char_p, vp]),
        "mde_context_set_input_shape": (C.c_int, [vp, C.c_char_p, i32, P(i64)]),
        "mde_context_enqueue": (C.c_int, [vp, vp]),
        "mde_context_set_gather": (C.c_int, [vp, i32, i32, P(vp)]),
        "mde_context_launches_per_enqueue": (C.c_int, [vp]),
        "mde_context_get_buffer": (C.c_int, [vp, C.c_char_p, P(vp), P(i64), P(i32)]),
        "mde_context_snapshot_block": (C.c_int, [vp, i32]),
        "mde_context_enqueue_timed": (C.c_int, [vp, vp, P(f32), i32]),
        "mde_context_op_info": (C.c_int, [vp, i32, C.c_char_p, i32, P(C.c_double), P(C.c_double)]),
        "mde_k_preprocess_u8": (C.c_int, [i32, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                          P(C.c_double), P(C.c_double), vp, vp, vp]),
        "mde_k_preprocess_u8_pad": (C.c_int, [i32, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                              P(C.c_double), P(C.c_double), P(C.c_double), vp, vp, vp]),
        "mde_k_preprocess_u8_square_pad_cubic": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp]),
        "mde_k_preprocess_u8_cubic_f32": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, P(C.c_double), P(C.c_double), vp, vp]),
        "mde_k_im2col_f32": (C.c_int, [i32, vp, i32, i32, i32, i32, i32, vp, vp]),
        "mde_k_gemm": (C.c_int, [i32, vp, i64, i32, i32, vp, i32, i32, P(Epilogue), vp]),
        "mde_k_gemm_tiled": (C.c_int, [i32, vp, i64, i32, i32, vp, i32, i32, P(Epilogue), i32, i32, i32, vp]),
        "mde_k_conv3x3": (C.c_int, [i32, vp, i32, i32, i32, i32, vp, i32, P(Epilogue), vp]),
        "mde_k_attention": (C.c_int, [i32, vp, vp, i32, i32, i32, vp]),
        "mde_k_attention_mma": (C.c_int, [i32, vp, vp, i32, i32, i32, vp]),
        "mde_k_attention_kv": (C.c_int, [i32, vp, i32, vp, i32, i32, i32, vp, i32, i32, i32, i32, vp]),
        "mde_k_attention_poly": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, vp]),
        "mde_k_attention_q3": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, vp]),
        "mde_k_attention_trace": (C.c_int, [i32, vp, vp, i32, i32, i32, vp, vp]),
        "mde_k_layernorm": (C.c_int, [i32, vp, vp, vp, vp, i64, i32, f32, i32, i32, vp]),
        "mde_k_bilinear": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
        "mde_k_bilinear_add": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp]),
        "mde_k_assemble_tokens": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, i32, vp, vp]),
        "mde_k_im2col_s2": (C.c_int, [i32, vp, vp, i32, i32, i32, i32, vp]),
        "mde_k_peer_signal": (C.c_int, [P(vp), i32, i32, C.c_uint32, vp]),
        "mde_k_peer_wait": (C.c_int, [vp, i32, C.c_uint32, vp]),
        "mde_k_peer_signal_counter": (C.c_int, [P(vp), i32, i32, vp, i32, vp]),
        "mde_k_peer_wait_counter": (C.c_int, [vp, i32, vp, vp]),
        "mde_k_merge_patches": (C.c_int, [i32, vp, i32, i32, i32, i32, vp, vp]),
        "mde_k_resize_depth": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, f32, f32, vp]),
        "mde_k_resize_crops": (C.c_int, [vp, i32, i32, i32, i32, P(Crop), i32, i32, i32, P(f32), P(f32), vp, vp]),
        "mde_k_qknorm_rope": (C.c_int, [i32, vp, i64, i32, vp, vp, vp, vp, f32, vp, vp, i32, i32, P(vp), i32, vp]),
        "mde_k_depth_pro_post": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp, vp]),
        "mde_k_resize_depth_halfpixel": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, f32, f32, f32, vp]),
        "mde_k_resize_depth_halfpixel_nan": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, f32, vp]),
        "mde_k_upconv_head": (C.c_int, [i32, vp, i32, i32, i32, i32, i32, i32, vp, vp, f32, f32, vp, vp]),
    }
    assert sorted(protos) == sorted(SYMBOLS)
    for name, (res, args) in protos.items():
        fn = getattr(lib, name)      # AttributeError here == the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().mde_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    """Non-zero return -> RuntimeError, like core/common_runtime.py:41-56 `cuda_call` in the reference."""
    if rc != 0:
        raise RuntimeError(f"[MDET] {what or 'libmde_b200'} failed (code {rc}): {last_error()}")
