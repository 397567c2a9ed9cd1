"""ctypes binding of liblipread_b200.so.  The signatures are read from include/lipread_b200.h (the
single source of truth for the C ABI), so the binding cannot drift from the header.  There is no
fallback: if the library is missing the import fails loudly."""
import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "liblipread_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "lipread_b200.h")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing. Build it with `python -m multimodal_lipread_b200.build` "
        "(nvcc, sm_100a). multimodal_lipread_b200 has no CPU / PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)

_SCALARS = {
    "int": ctypes.c_int, "float": ctypes.c_float, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
    "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong, "lr_stream_t": ctypes.c_void_p,
}


def _ctype(decl, is_return=False):
    decl = re.sub(r"\bconst\b", "", decl).strip()
    if "*" in decl:
        return ctypes.c_char_p if (is_return and decl.startswith("char")) else ctypes.c_void_p
    words = decl.split()
    if not is_return:
        words = words[:-1]               # every parameter in the header is named
    key = " ".join(words)
    if key not in _SCALARS:
        raise ValueError(f"cannot map C declaration {decl!r}")
    return _SCALARS[key]


def parse_header(path=HEADER_PATH):
    """name -> (restype, [argtypes]) for every `lr_*` function declared in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    sigs = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(lr_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        argtypes = [] if args in ("", "void") else [_ctype(a) for a in args.split(",")]
        sigs[name] = (_ctype(ret, is_return=True), argtypes)
    return sigs


SIGNATURES = parse_header()

LR_LOGMEL_FRONTEND, LR_LOGMEL_RAW = 0, 1
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_RELU6 = 0, 1, 2, 3, 4


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
