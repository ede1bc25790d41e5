"""ctypes binding of liboasr.so (C-ABI declared in include/oasr.h).

There is no CPU fallback: if the shared library is missing, or a call fails, this module raises.
Error mapping follows the reference's exception discipline (gemini_pipeline.py:31-35, 329-334, 739-741):
OASR_ERR_INVALID / OASR_ERR_UNSUPPORTED -> ValueError, everything else -> RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "_lib" / "liboasr.so"

OASR_OK = 0
ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_UNSUPPORTED = -1, -2, -3, -4
DTYPE_F32, DTYPE_BF16 = 0, 1
EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_F32_RESID, EPI_ARGMAX, EPI_LN_GELU_BF16, EPI_F32_GELU_RESID = range(7)
FLAG_INPUT_NORMALISED = 1
FLAG_INPUT_I16 = 2
PROF_CATEGORIES = ("wave_norm", "fe_layer0", "fe_conv_1_6", "layernorm", "feature_proj", "posconv", "qkv_gemm",
                   "attention", "outproj_gemm", "ffn1_gemm", "ffn2_gemm", "ctc_head_argmax", "decode", "tp_allreduce", "end")


class OasrConfig(C.Structure):
    _fields_ = [
        ("d_model", C.c_int32), ("n_layers", C.c_int32), ("n_heads", C.c_int32), ("d_ffn", C.c_int32),
        ("vocab", C.c_int32), ("fe_dim", C.c_int32), ("pos_kernel", C.c_int32), ("pos_groups", C.c_int32),
        ("n_fe_layers", C.c_int32), ("fe_kernel", C.c_int32 * 8), ("fe_stride", C.c_int32 * 8),
        ("blank_id", C.c_int32),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_PROTOTYPES = {
    # name: (restype, [argtypes])
    "oasr_version": (C.c_char_p, []),
    "oasr_last_error": (C.c_char_p, []),
    "oasr_create": (C.c_int, [C.POINTER(OasrConfig), C.POINTER(_vp)]),
    "oasr_destroy": (None, [_vp]),
    "oasr_load_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int, C.POINTER(_i64), C.c_int]),
    "oasr_finalize_weights": (C.c_int, [_vp]),
    "oasr_tp_unique_id": (C.c_int, [_vp]),
    "oasr_tp_init": (C.c_int, [_vp, _i32, _i32, _vp]),
    "oasr_tp_emulate": (C.c_int, [_vp, _i32]),
    "oasr_tp_ipc_export": (C.c_int, [_vp, _i32, _i32, _vp]),
    "oasr_tp_ipc_import": (C.c_int, [_vp, _vp]),
    "oasr_feature_length": (_i32, [C.POINTER(OasrConfig), _i64]),
    "oasr_forward_ctc": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "oasr_transcribe_host": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "oasr_transcribe_host_async": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp,
                                             C.POINTER(_i64)]),
    "oasr_wait": (C.c_int, [_vp, _i64]),
    "oasr_debug_forward": (C.c_int, [_vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _vp]),
    "oasr_debug_buffer": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i32)]),
    "oasr_debug_copy": (C.c_int, [_vp, C.c_char_p, _vp, _i64]),
    "oasr_launch_count": (_i64, [_vp]),
    "oasr_profile_enable": (C.c_int, [_vp, _i32]),
    "oasr_profile_read": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(_i64), _i32]),
    "oasr_wave_norm": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "oasr_resample_length": (_i64, [_i64, _i32, _i32]),
    "oasr_resample": (C.c_int, [_vp, _i32, _i64, _i32, _i32, _i32, _vp, _i64, _vp]),
    "oasr_fe_layer0": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "oasr_conv_ln_gelu": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "oasr_layernorm": (C.c_int, [_vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp]),
    "oasr_gemm": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    "oasr_posconv": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "oasr_attention": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp]),
    "oasr_ctc_decode": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "oasr_ctc_collapse": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("OASR_LIB", str(_LIB_PATH)))


def load() -> C.CDLL:
    """Load liboasr.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise RuntimeError(
            f"liboasr.so not found at {path}. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C omnilingual-asr_b200/csrc`). There is no CPU fallback for the CTC path.")
    lib = C.CDLL(str(path))
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().oasr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "liboasr") -> None:
    if rc == OASR_OK:
        return
    msg = f"{what} failed ({rc}): {last_error()}"
    if rc in (ERR_INVALID, ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)


def ptr(t) -> C.c_void_p:
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return C.c_void_p(0)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def stream_ptr(stream=None) -> C.c_void_p:
    if stream is None:
        import torch
        stream = torch.cuda.current_stream()
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream) or 0))
