"""ctypes binding of libdetsam2.so (the C ABI declared in include/detsam2.h).

The library is the only compute path of the product: if it cannot be loaded, importing the
kernels raises — there is no CPU or eager fallback.
"""
import ctypes as C
import os

from . import build as _build

_LIB = None


class Ds2Error(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p),
        ("lda", C.c_int64), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", C.c_void_p), ("gamma", C.c_void_p), ("residual", C.c_void_p),
        ("ldr", C.c_int64),
        ("res_row_mod", C.c_int32), ("act", C.c_int32),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p),
        ("ldc", C.c_int64), ("ldc_bf16", C.c_int64),
        ("rope_cs", C.c_void_p),
        ("rope_col0", C.c_int32), ("rope_col1", C.c_int32),
        ("rope_period", C.c_int32), ("rope_rows_per_batch", C.c_int32),
        ("rope_row_limit", C.c_int32), ("impl", C.c_int32),
        ("rope_axial", C.c_void_p), ("rope_side", C.c_int32),
    ]


class FlashArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("ldq", C.c_int64), ("ldk", C.c_int64), ("ldv", C.c_int64), ("ldo", C.c_int64),
        ("bsq", C.c_int64), ("bsk", C.c_int64), ("bsv", C.c_int64), ("bso", C.c_int64),
        ("B", C.c_int32), ("Lq", C.c_int32), ("Lk", C.c_int32), ("DV", C.c_int32),
        ("scale", C.c_float), ("impl", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("impl_flags", C.c_int32),
    ]


class MhaArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("out", C.c_void_p),
        ("q_tok_stride", C.c_int64), ("k_tok_stride", C.c_int64),
        ("v_tok_stride", C.c_int64), ("o_tok_stride", C.c_int64),
        ("q_bs", C.c_int64), ("k_bs", C.c_int64), ("v_bs", C.c_int64), ("o_bs", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("D", C.c_int32),
        ("Lq", C.c_int32), ("Lk", C.c_int32),
        ("window", C.c_int32), ("Hm", C.c_int32), ("Wm", C.c_int32), ("q_pool", C.c_int32),
        ("Lk_valid", C.c_int32),
        ("scale", C.c_float),
        ("pad_q", C.c_void_p), ("pad_k", C.c_void_p), ("pad_v", C.c_void_p),
    ]


class LnArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("rows", C.c_int32), ("C", C.c_int32),
        ("w", C.c_void_p), ("b", C.c_void_p), ("eps", C.c_float),
        ("act", C.c_int32),
        ("out_f32", C.c_void_p), ("out_bf16", C.c_void_p), ("ldo", C.c_int64),
        ("pos", C.c_void_p), ("pos_row_mod", C.c_int32),
        ("out2_bf16", C.c_void_p),
        ("x_bf16", C.c_void_p),
    ]


class Mlp3Args(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int64), ("gather", C.c_void_p),
        ("rows", C.c_int32), ("nmlp", C.c_int32), ("din", C.c_int32), ("dh", C.c_int32),
        ("dout", C.c_int32),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("sigmoid_out", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int64),
    ]


# every exported symbol of include/detsam2.h: name -> (restype, argtypes)
_P, _I, _L, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SYMBOLS = {
    "ds2_version": (C.c_int, []),
    "ds2_launch_count": (C.c_int64, []),
    "ds2_last_error": (C.c_char_p, []),
    "ds2_device_sm_count": (C.c_int, []),
    "ds2_gemm": (C.c_int, [C.POINTER(GemmArgs), _P]),
    "ds2_flash_attn": (C.c_int, [C.POINTER(FlashArgs), _P]),
    "ds2_flash_workspace_bytes": (C.c_int64, [_I, _I, _I]),
    "ds2_debug_flash_stalls": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int]),
    "ds2_debug_win_times": (C.c_int, [C.POINTER(C.c_longlong)]),
    "ds2_mha": (C.c_int, [C.POINTER(MhaArgs), _P]),
    "ds2_layernorm": (C.c_int, [C.POINTER(LnArgs), _P]),
    "ds2_axpby": (C.c_int, [_P, _P, _L, _I, _I, _F, _F, _P, _P, _P]),
    "ds2_cast_f32_bf16": (C.c_int, [_P, _P, _L, _P]),
    "ds2_cast_bf16_f32": (C.c_int, [_P, _P, _L, _P]),
    "ds2_maxpool2x2": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "ds2_upsample2x_add": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "ds2_im2col_patch": (C.c_int, [_P, _P, _I, _I, _P]),
    "ds2_im2col_k3s2": (C.c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "ds2_dwconv7": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ds2_maskds_stage1": (C.c_int, [_P, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P]),
    "ds2_maskds_conv": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "ds2_upscale1": (C.c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "ds2_upscale2_masks": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "ds2_mlp3": (C.c_int, [C.POINTER(Mlp3Args), _P]),
    "ds2_sam_select": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P]),
    "ds2_objptr_mix": (C.c_int, [_P, _P, _P, _I, _I, _P]),
    "ds2_bank_gather": (C.c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P]),
    "ds2_bank_ptr": (C.c_int, [_P, _P, _P, _P, _I, _L, _I, _P]),
    "ds2_prompt_tokens": (C.c_int, [_P, _P, _I, _I, _P, _P, _P, _P, _I, _F, _P, _P]),
    "ds2_bank_ptr_pe": (C.c_int, [_P, _F, _P, _P, _P, _P, _I, _L, _I, _P]),
    "ds2_bank_assemble": (C.c_int, [_P, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "ds2_memenc_finish": (C.c_int, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "ds2_connected_components": (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "ds2_fill_holes": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "ds2_resize_bilinear": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "ds2_threshold_pack": (C.c_int, [_P, _P, _L, _P]),
    "ds2_mask_pack_stats": (C.c_int, [_P, _P, _P, _I, _I, _I, _P]),
    "ds2_downsample4_aa": (C.c_int, [_P, _P, _I, _I, _F, _F, _P]),
    "ds2_mask_prompt_embed": (C.c_int, [_P, _I, _I] + [_P] * 10 + [_P, _P]),
    "ds2_ingest_frames": (C.c_int, [_P, _I, _I, _I, _L, _L, _P, _P, _I, _P]),
    "ds2_letterbox_frames": (C.c_int, [_P, _I, _I, _I, _L, _L, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
}


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Loads libdetsam2.so (building it with nvcc when absent) and types every symbol."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise Ds2Error(f"{path} not built; run `python -m detsam2_b200.build`")
        _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # raises AttributeError when the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().ds2_last_error()
        raise Ds2Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
