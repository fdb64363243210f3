"""ctypes binding of libavmoe_b200.so (the C ABI declared in include/avmoe_b200.h).

The library is the ONLY compute path: if it is missing or a call fails, a RuntimeError is
raised — there is no eager / CPU fallback (BASELINE.json north_star)."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("B", C.c_void_p), ("D", C.c_void_p),
        ("dtype", C.c_int32), ("d_dtype", C.c_int32),
        ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
        ("a_kmajor", C.c_int32), ("b_kmajor", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64), ("ldd", C.c_int64),
        ("a_batch", C.c_int32), ("b_batch", C.c_int32),
        ("a_bstride", C.c_int64), ("b_bstride", C.c_int64), ("d_bstride", C.c_int64),
        ("Nb", C.c_int64),
        ("alpha", C.c_float),
        ("row_bias", C.c_void_p), ("col_scale", C.c_void_p), ("col_bias", C.c_void_p),
        ("split_k", C.c_int32), ("accumulate", C.c_int32), ("impl", C.c_int32),
    ]


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Load the shared library (building it in-tree first if sources are newer)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    if not os.path.exists(path) or (os.environ.get("AVMOE_AUTOBUILD", "1") == "1" and _build.needs_build()):
        try:
            _build.build()
        except Exception as e:  # no nvcc on the box and no prebuilt .so
            if not os.path.exists(path):
                raise RuntimeError(f"avmoe_b200: CUDA extension missing and could not be built: {e}") from e
    try:
        lib = C.CDLL(path)
    except OSError as e:
        raise RuntimeError(f"avmoe_b200: cannot load {path}: {e}") from e
    lib.avmoe_abi_version.restype = C.c_int
    lib.avmoe_build_info.restype = C.c_char_p
    lib.avmoe_last_error.restype = C.c_char_p
    _LIB = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().avmoe_last_error().decode(errors="replace")
        raise RuntimeError(f"avmoe_b200.{what} failed (status {rc}): {msg}")


def exported_symbols():
    """Names declared in include/avmoe_b200.h (parsed), for the ABI export test."""
    import re
    hdr = os.path.join(os.path.dirname(_build.PKG_DIR), "include", "avmoe_b200.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(avmoe_[a-z0-9_]+)\s*\(", txt)))
