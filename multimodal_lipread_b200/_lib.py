"""ctypes binding of liblipread_b200.so (include/lipread_b200.h).  There is no fallback: if the
library is missing the import fails loudly."""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblipread_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python -m multimodal_lipread_b200.build` "
        "(nvcc, sm_100a). multimodal_lipread_b200 has no CPU / PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)

c_int, c_size_t, c_void_p, c_float, c_double = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_float, ctypes.c_double
c_ll = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/lipread_b200.h one to one
SIGNATURES = {
    "lr_version": (c_int, []),
    "lr_last_error": (ctypes.c_char_p, []),
    "lr_launch_count": (ctypes.c_ulonglong, []),
    "lr_logmel_plan_bytes": (c_size_t, []),
    "lr_logmel_plan_init": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lr_logmel_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lr_normalize_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
}

LR_LOGMEL_FRONTEND, LR_LOGMEL_RAW = 0, 1


def _bind():
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args


_bind()


class LipreadError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise LipreadError(f"lipread_b200 error {rc}: {lib.lr_last_error().decode()}")


def launch_count():
    return int(lib.lr_launch_count())
