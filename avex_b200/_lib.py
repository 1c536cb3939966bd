"""ctypes binding of libavexk.so (the C ABI declared in include/avexk.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavexk.so")


class AvexkError(RuntimeError):
    pass


class BeatsDims(C.Structure):
    _fields_ = [
        ("layers", C.c_int),
        ("embed", C.c_int),
        ("ffn", C.c_int),
        ("heads", C.c_int),
        ("patch_embed", C.c_int),
        ("conv_pos", C.c_int),
        ("conv_groups", C.c_int),
        ("fbank_mean", C.c_float),
        ("fbank_std", C.c_float),
        ("ln_eps", C.c_float),
    ]


_LAYER_FIELDS = [
    "q_w", "q_b", "k_w", "k_b", "v_w", "v_b", "o_w", "o_b",
    "grep_w", "grep_b", "grep_a",
    "ln1_w", "ln1_b",
    "fc1_w", "fc1_b", "fc2_w", "fc2_b",
    "ln2_w", "ln2_b",
]  # fmt: skip


class BeatsLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


_TOP_FIELDS = [
    "patch_w", "ln0_w", "ln0_b", "proj_w", "proj_b",
    "posconv_g", "posconv_v", "posconv_b", "enc_ln_w", "enc_ln_b", "rel_bias_table",
]  # fmt: skip


class BeatsWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _TOP_FIELDS] + [("layers", C.POINTER(BeatsLayerWeights))]


# name -> (restype, argtypes); every symbol include/avexk.h declares
_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
SIGNATURES = {
    "avexk_last_error": (C.c_char_p, []),
    "avexk_version": (_i, []),
    "avexk_build_id": (C.c_char_p, []),
    "avexk_launch_count": (_ll, []),
    "avexk_profile_enable": (None, [_i]),
    "avexk_profile_read": (_i, [_i, C.POINTER(_ll), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "avexk_fbank_create": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "avexk_fbank_destroy": (None, [_vp]),
    "avexk_fbank_num_frames": (_i, [_i]),
    "avexk_fbank_forward": (_i, [_vp, _vp, _i, _i, _ll, _f, _f, _f, _i, _i, _vp, _vp, _i, _vp]),
    "avexk_fbank_patch_operand": (_i, [_vp, _vp, _i, _i, _ll, _f, _f, _f, _vp, _vp]),
    "avexk_gemm_bf16": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _i, _vp, _vp, _f, _vp, _ll, _i, _vp]),
    "avexk_beats_set_precision": (_i, [_vp, _i]),
    "avexk_gemm_ln_scratch_bytes": (_sz, [_i]),
    "avexk_gemm_bf16_ln": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _f, _vp, _vp, _vp, _sz, _vp]),
    "avexk_gemm_bf16_ln_pooled": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _f, _vp, _vp, _vp, _sz, _i, _vp, _vp, _vp, _vp]),
    "avexk_gemm_config": (_i, [_i]),
    "avexk_layernorm": (_i, [_vp, _i, _i, _vp, _vp, _f, _vp, _vp, _vp]),
    "avexk_attention_gated": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "avexk_posconv_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "avexk_posconv": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "avexk_beats_create": (_i, [C.POINTER(BeatsDims), C.POINTER(_vp)]),
    "avexk_beats_destroy": (None, [_vp]),
    "avexk_beats_load_weights": (_i, [_vp, C.POINTER(BeatsWeights), _vp]),
    "avexk_beats_num_tokens": (_i, [_i]),
    "avexk_beats_workspace_bytes": (_sz, [_vp, _i, _i]),
    "avexk_beats_forward": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp, _vp, _vp, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _sz, _vp]),
}



class EffnetBlockCfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("kernel", "stride", "cin", "cexp", "cout", "csq")]


class BnParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("weight", "bias", "mean", "var")]


class EffnetBlockWeights(C.Structure):
    _fields_ = [
        ("expand_w", C.c_void_p), ("expand_bn", BnParams),
        ("dw_w", C.c_void_p), ("dw_bn", BnParams),
        ("se1_w", C.c_void_p), ("se1_b", C.c_void_p), ("se2_w", C.c_void_p), ("se2_b", C.c_void_p),
        ("proj_w", C.c_void_p), ("proj_bn", BnParams),
    ]  # fmt: skip


class EffnetWeights(C.Structure):
    _fields_ = [
        ("stem_w", C.c_void_p), ("stem_bn", BnParams),
        ("blocks", C.POINTER(EffnetBlockWeights)),
        ("head_w", C.c_void_p), ("head_bn", BnParams),
        ("cls_w", C.c_void_p), ("cls_b", C.c_void_p), ("num_classes", C.c_int),
    ]  # fmt: skip


SIGNATURES.update({
    "avexk_melspec_create": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "avexk_melspec_destroy": (None, [_vp]),
    "avexk_melspec_num_frames": (_i, [_i]),
    "avexk_melspec_forward": (_i, [_vp, _vp, _i, _i, _ll, _i, _vp, _vp, _vp]),
    "avexk_conv1x1_f16": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "avexk_conv1x1_se_f16": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "avexk_dwconv_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "avexk_effnet_create": (_i, [C.POINTER(EffnetBlockCfg), _i, _i, _i, C.POINTER(_vp)]),
    "avexk_effnet_destroy": (None, [_vp]),
    "avexk_effnet_load_weights": (_i, [_vp, C.POINTER(EffnetWeights), _vp]),
    "avexk_effnet_out_hw": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "avexk_effnet_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "avexk_effnet_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, C.POINTER(_vp), _vp, _sz, _vp]),
})

_lib = None


def load() -> C.CDLL:
    """Load libavexk.so (built in-tree by `python -m avex_b200.build`).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AvexkError(
            f"{LIB_PATH} not found: build it with `python -m avex_b200.build` (needs nvcc). "
            "avex_b200 has no CPU or eager fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    # a stale binary must never be what runs: the .so is git-ignored and travels with the snapshot, so compare the id compiled
    # into it with the hash of the sources lying next to it (skipped only when the sources are not shipped at all)
    from . import build as _build

    if os.path.isdir(_build.CSRC):
        want, got = _build.source_id(), lib.avexk_build_id().decode()
        if want != got:
            raise AvexkError(
                f"{LIB_PATH} is stale: built from sources {got}, the tree hashes to {want}. "
                "Rebuild with `python -m avex_b200.build`."
            )
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().avexk_last_error().decode("utf-8", "replace")
        raise AvexkError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().avexk_launch_count())
