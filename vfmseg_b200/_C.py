"""ctypes binding of libvfmseg_b200.so (the C ABI declared in include/vfmseg_b200.h).

There is no fallback: if the shared library is missing or a call fails, this raises. The product
path never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from .build import LIB_PATH

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ll = C.c_longlong
_sz = C.c_size_t


class VfmPixelNorm(C.Structure):
    _fields_ = [("mean", _f * 3), ("inv_std", _f * 3), ("flip", _i)]


class VfmBlockParams(C.Structure):
    _fields_ = [(n, _p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ls1",
        "ln2_w", "ln2_b", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2",
        "qkv_wf", "qkv_bf", "qkv_cs", "fc1_wf", "fc1_bf", "fc1_cs")]


class VfmVitParams(C.Structure):
    _fields_ = [
        ("embed_dim", _i), ("depth", _i), ("heads", _i), ("mlp_hidden", _i), ("n_taps", _i),
        ("tap_blocks", _i * 8), ("ln_eps", _f),
        ("patch_w", _p), ("patch_b", _p), ("cls_token", _p), ("pos_embed", _p),
        ("blocks", C.POINTER(VfmBlockParams)),
    ]


class VfmEvaBlockParams(C.Structure):
    _fields_ = [(n, _p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "w12", "b12",
        "ffn_ln_w", "ffn_ln_b", "w3", "b3", "qkv_wf", "qkv_bf", "qkv_cs", "w12_wf", "w12_bf", "w12_cs")]


class VfmEvaParams(C.Structure):
    _fields_ = [
        ("embed_dim", _i), ("depth", _i), ("heads", _i), ("hidden", _i), ("hidden_pad", _i), ("n_taps", _i), ("grid", _i),
        ("tap_blocks", _i * 8), ("ln_eps", _f),
        ("patch_w", _p), ("patch_b", _p), ("cls_token", _p), ("pos_embed", _p), ("rope_cos", _p), ("rope_sin", _p), ("ones", _p),
        ("blocks", C.POINTER(VfmEvaBlockParams)),
    ]


class VfmSamBlockParams(C.Structure):
    _fields_ = [("window", _i), ("qkv_n", _i)] + [(n, _p) for n in (
        "ln1_w", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "lin1_w", "lin1_b", "lin2_w", "lin2_b",
        "lin1_wf", "lin1_bf", "lin1_cs")]


class VfmSamParams(C.Structure):
    _fields_ = [
        ("embed_dim", _i), ("depth", _i), ("heads", _i), ("head_dim", _i), ("hidden", _i), ("n_taps", _i), ("grid", _i),
        ("use_rel_pos", _i), ("tap_blocks", _i * 8), ("ln_eps", _f),
        ("patch_w", _p), ("patch_b", _p), ("pos_embed", _p), ("ones", _p),
        ("part", _p), ("unpart", _p), ("win_rows", _i), ("win_buf", _p),
        ("onehot", _p), ("onehot_rows", _i),
        ("blocks", C.POINTER(VfmSamBlockParams)),
    ]


class VfmLinearHeadParams(C.Structure):
    _fields_ = [
        ("in_channels", _i), ("mid_channels", _i), ("groups", _i), ("num_classes", _i), ("gn_eps", _f),
        ("fusion_w", _p), ("gn_w", _p), ("gn_b", _p),
        ("up1_w", _p), ("up1_b", _p), ("up2_w", _p), ("up2_b", _p), ("cls_w", _p), ("cls_b", _p),
    ]


# name -> (restype, argtypes); must list every symbol include/vfmseg_b200.h declares
SIGNATURES = {
    "vfm_last_error": (C.c_char_p, []),
    "vfm_abi_version": (_i, []),
    "vfm_launch_count": (_ll, []),
    "vfm_device_check": (_i, []),
    "vfm_prof_enable": (_i, [_i]),
    "vfm_prof_report": (_i, [C.c_char_p, _sz]),
    "vfm_gemm_bias_bf16": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "vfm_gemm_bias_rope_bf16": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p]),
    "vfm_gemm_bias_gelu_bf16": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "vfm_gemm_bias_ls_residual": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vfm_gemm_bias_ls_residual_stats": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p]),
    "vfm_gemm_lnfold_bf16": (_i, [_p, _i, _p, _i, _p, _p, _p, _f, _i, _p, _i, _i, _i, _i, _p]),
    "vfm_gemm_lnfold_rope_bf16": (_i, [_p, _i, _p, _i, _p, _p, _p, _f, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p]),
    "vfm_gemm_patch_embed": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _p]),
    "vfm_gemm_convt2x2_gelu": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vfm_gemm_cls_nchw": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "vfm_gemm_f32": (_i, [_p, _i, _p, _i, _p, _p, _i, _i, _i, _i, _p]),
    "vfm_attention_fwd": (_i, [_p, _p, _i, _i, _i, _p]),
    "vfm_attention_fwd_ex": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vfm_attention_cross": (_i, [_p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p]),
    "vfm_image_resize_norm": (_i, [_p, _i, C.POINTER(VfmPixelNorm), _i, _i, _i, _p, _i, _i, _p]),
    "vfm_ms_confidence": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _p, _p]),
    "vfm_ms_context_im2col": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p]),
    "vfm_space_to_depth2": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vfm_groupnorm_act": (_i, [_p, _p, _i, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "vfm_geglu": (_i, [_p, _p, _ll, _i, _p]),
    "vfm_cast_f32_bf16": (_i, [_p, _p, _ll, _p]),
    "vfm_ms_merge_argmax": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "vfm_rope_qk": (_i, [_p, _ll, _i, _i, _i, _p, _p, _p]),
    "vfm_swiglu_layernorm": (_i, [_p, _p, _p, _p, _ll, _i, _i, _f, _p]),
    "vfm_patch_gather": (_i, [_p, _i, C.POINTER(VfmPixelNorm), _i, _i, _p, _i, _i, _i, _p, _p]),
    "vfm_cls_rows": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "vfm_layernorm": (_i, [_p, _p, _p, _p, _i, _i, _f, _p]),
    "vfm_layernorm_tap": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _i, _i, _i, _p]),
    "vfm_gemm_patch_embed_ex": (_i, [_p, _i, _p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "vfm_layernorm_tap_ex": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _i, _i, _i, _i, _p]),
    "vfm_relpos_terms": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "vfm_rows_gather": (_i, [_p, _p, _p, _ll, _i, _p]),
    "vfm_attention_relpos": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "vfm_attention_global_tc": (_i, [_p, _i, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "vfm_attention_window_tc_map": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "vfm_layernorm_tap_map": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _i, _i, _i, _i, _p, _p]),
    "vfm_attention_window_tc": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "vfm_attention_relpos_ex": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _i, _f, _p]),
    "vfm_groupnorm_relu": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f, _i, _p]),
    "vfm_slide_merge_argmax": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "vfm_confusion_matrix": (_i, [_p, _p, _ll, _i, _i, _p, _p]),
    "vfm_slide_merge_flip_argmax": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vfm_tta_flip_mean_argmax": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "vfm_vit_workspace_bytes": (_sz, [C.POINTER(VfmVitParams), _i, _i, _i]),
    "vfm_vit_forward": (_i, [C.POINTER(VfmVitParams), _p, _i, C.POINTER(VfmPixelNorm), _i, _i, _p, _i, _i, _i,
                             _p, _p, _sz, _p]),
    "vfm_eva_workspace_bytes": (_sz, [C.POINTER(VfmEvaParams), _i]),
    "vfm_eva_forward": (_i, [C.POINTER(VfmEvaParams), _p, _i, C.POINTER(VfmPixelNorm), _i, _i, _p, _i, _p, _p, _sz, _p]),
    "vfm_sam_workspace_bytes": (_sz, [C.POINTER(VfmSamParams), _i]),
    "vfm_sam_forward": (_i, [C.POINTER(VfmSamParams), _p, _i, C.POINTER(VfmPixelNorm), _i, _i, _p, _i, _p, _p, _sz, _p]),
    "vfm_linear_head_workspace_bytes": (_sz, [C.POINTER(VfmLinearHeadParams), _i, _i, _i]),
    "vfm_linear_head_forward": (_i, [C.POINTER(VfmLinearHeadParams), _p, _i, _i, _i, _p, _p, _sz, _p]),
}


class VfmError(RuntimeError):
    pass


_lib = None


def lib_path() -> Path:
    return LIB_PATH


def load():
    """Load (once) and type the shared library. Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VfmError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). vfmseg_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and this table diverge
        fn.restype = res
        fn.argtypes = args
    if lib.vfm_abi_version() != 2:
        raise VfmError(f"ABI version mismatch: library {lib.vfm_abi_version()}, binding 2")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().vfm_last_error()
        raise VfmError(f"vfmseg_b200 error {rc}: {msg.decode() if msg else '?'}")


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero status."""
    check(getattr(load(), name)(*args))
