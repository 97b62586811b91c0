"""ctypes binding of libuce_b200.so (include/uce_b200.h).  No fallback: if the library is
missing or the device is not a B200-class GPU every call raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libuce_b200.so")

UCE_E_ARG, UCE_E_STATE, UCE_E_NOT_SPD, UCE_E_NO_DEVICE = -1, -2, -3, -4

_lib = None

# name -> (restype, argtypes); mirrors include/uce_b200.h one to one
_PP_F = C.POINTER(C.c_void_p)
SIGNATURES = {
    "uce_abi_version": (C.c_int, []),
    "uce_last_error": (C.c_char_p, []),
    "uce_ws_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "uce_ws_destroy": (C.c_int, [C.c_void_p]),
    "uce_ws_set_apply_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "uce_plan_row_blocks": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "uce_ws_set_factor_impl": (C.c_int, [C.c_void_p, C.c_int]),
    "uce_ws_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "uce_ws_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "uce_ws_timings": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "uce_factor_dev_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "uce_apply_dev_f32": (C.c_int, [C.c_void_p, _PP_F, _PP_F, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "uce_edit_dev_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float,
                                   _PP_F, _PP_F, C.POINTER(C.c_int), C.c_int, C.c_void_p]),
    "uce_edit_host_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_float,
                                    _PP_F, _PP_F, C.POINTER(C.c_int), C.c_int]),
    "uce_ws_check": (C.c_int, [C.c_void_p, C.c_void_p]),
    "uce_ws_info": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 6),
    "uce_ws_debug_read": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "uce_artifact_write_f32": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), _PP_F, C.POINTER(C.c_long), C.POINTER(C.c_long)]),
    "uce_artifact_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "uce_artifact_count": (C.c_int, [C.c_void_p]),
    "uce_artifact_entry": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_long)]),
    "uce_artifact_read_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "uce_artifact_close": (C.c_int, [C.c_void_p]),
    "uce_png_write_rgb8": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
}


class UCEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libuce_b200 error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                               "(there is no CPU fallback for the UCE hot path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise UCEError(rc, lib().uce_last_error().decode(errors="replace"))
