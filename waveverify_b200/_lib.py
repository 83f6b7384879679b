"""ctypes binding of libwv_b200.so (include/wv_b200.h).  No fallback: if the CUDA library is
missing or fails to load, importing a compute entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WV_LIB_PATH") or os.path.join(HERE, "libwv_b200.so")   # override: A/B builds

KIND = {"generator": 0, "detector": 1, "locator": 2}


class NetConfigC(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("sample_rate", C.c_int), ("dimension", C.c_int),
        ("channels_enc", C.c_int), ("channels_dec", C.c_int), ("n_fft_base", C.c_int),
        ("n_residual_enc", C.c_int), ("n_residual_dec", C.c_int), ("n_strides", C.c_int),
        ("strides", C.c_int * 8), ("res_scale_enc", C.c_float), ("res_scale_dec", C.c_float),
        ("nbits", C.c_int), ("output_dim", C.c_int), ("msg_dimension", C.c_int),
        ("embedding_dim", C.c_int), ("embedding_layers", C.c_int), ("freq_bands", C.c_int),
        ("precise", C.c_int),
    ]


class TensorC(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int),
                ("shape", C.c_int64 * 4)]


_lib = None

_SIGS = {
    "wv_version": (C.c_int, []),
    "wv_last_error": (C.c_char_p, []),
    "wv_net_create": (C.c_int, [C.POINTER(NetConfigC), C.POINTER(TensorC), C.c_int, C.c_int,
                                C.POINTER(C.c_void_p)]),
    "wv_net_destroy": (C.c_int, [C.c_void_p]),
    "wv_net_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "wv_net_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "wv_net_launches": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "wv_net_set_chunk": (C.c_int, [C.c_void_p, C.c_int]),
    "wv_generator_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wv_generator_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]),
    "wv_generator_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "wv_detector_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "wv_detector_refine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "wv_locator_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "wv_metrics_accumulate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p,
                                        C.c_void_p]),
    "wv_augment_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "wv_effect_pointwise": (C.c_int, [C.c_int, C.c_void_p, C.c_longlong, C.c_float, C.c_void_p, C.c_ulonglong,
                                      C.c_void_p, C.c_void_p]),
    "wv_effect_suppress": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wv_effect_median": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wv_effect_fir": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "wv_effect_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p]),
    "wv_net_set_range_check": (C.c_int, [C.c_void_p, C.c_int]),
    "wv_net_range_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]),
    "wv_net_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "wv_net_profile_read": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "wv_net_profile_tag": (C.c_char_p, [C.c_void_p, C.c_int]),
    "wv_debug_tap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int,
                               C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "wv_op_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                             C.c_void_p]),
    "wv_op_resblock": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "wv_op_gemm_dw5": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "wv_op_dw5": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                            C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "wv_op_down": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                             C.c_void_p]),
    "wv_op_up": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                           C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGS)


def lib() -> C.CDLL:
    """Load the CUDA library; raise loudly when it is missing (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m waveverify_b200.build` "
            "(waveverify_b200 has no CPU or PyTorch fallback)")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(l, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = lib().wv_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg} (code {rc})")
