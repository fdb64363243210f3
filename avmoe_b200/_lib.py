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
        ("split_k", C.c_int32), ("accumulate", C.c_int32), ("impl", C.c_int32), ("full_batch", C.c_int32),
        ("col_bias_bstride", C.c_int64),
    ]


class ResampleArgs(C.Structure):
    """avmoe_resample_args (include/avmoe_b200.h): the op-level token-resampler entry."""
    _fields_ = [
        ("dtype", C.c_int32), ("BT", C.c_int32), ("N", C.c_int32), ("Nx", C.c_int32), ("C", C.c_int32), ("Cx", C.c_int32),
        ("gemm_impl", C.c_int32),
        ("o", C.c_void_p), ("w_tok", C.c_void_p), ("b_tok", C.c_void_p), ("w_fc", C.c_void_p), ("b_fc", C.c_void_p),
        ("v", C.c_void_p), ("mid", C.c_void_p), ("s", C.c_void_p),
        ("dv", C.c_void_p), ("d_o", C.c_void_p), ("dw_tok", C.c_void_p), ("db_tok", C.c_void_p), ("dw_fc", C.c_void_p), ("db_fc", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64), ("launches", C.c_int32),
    ]


class RfTokArgs(C.Structure):
    """avmoe_rf_tok_args (include/avmoe_b200.h): row-fused passes of the latent-token update."""
    _fields_ = [
        ("BT", C.c_int32), ("Nk", C.c_int32), ("C", C.c_int32), ("T", C.c_int32), ("Tp", C.c_int32), ("per_sample", C.c_int32),
        ("kv", C.c_void_p), ("pack", C.c_void_p), ("part", C.c_void_p), ("A12", C.c_void_p), ("Dt", C.c_void_p), ("lse", C.c_void_p),
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


MAX_EXPERTS = 8


class Expert(C.Structure):
    _fields_ = [
        ("pre_kind", C.c_int32), ("relu", C.c_int32), ("latent_slot", C.c_int32), ("_pad", C.c_int32),
        ("gate_pre", C.c_void_p), ("aux", C.c_void_p), ("ln1_w", C.c_void_p), ("ln1_b", C.c_void_p),
        ("wd", C.c_void_p), ("wu", C.c_void_p), ("ln2_w", C.c_void_p), ("ln2_b", C.c_void_p), ("gate", C.c_void_p),
        ("g_gate_pre", C.c_void_p), ("g_aux", C.c_void_p), ("g_ln1_w", C.c_void_p), ("g_ln1_b", C.c_void_p),
        ("g_wd", C.c_void_p), ("g_wu", C.c_void_p), ("g_ln2_w", C.c_void_p), ("g_ln2_b", C.c_void_p), ("g_gate", C.c_void_p),
    ]


class ExpertArgs(C.Structure):
    _fields_ = [
        ("BT", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("Cr", C.c_int32), ("G", C.c_int32), ("T", C.c_int32),
        ("E", C.c_int32), ("nl", C.c_int32), ("dtype", C.c_int32), ("bn_mode", C.c_int32),
        ("ex", Expert * MAX_EXPERTS),
        ("x", C.c_void_p), ("tokp", C.c_void_p), ("z", C.c_void_p), ("stats1", C.c_void_p),
        ("aff1", C.c_void_p), ("save1", C.c_void_p), ("aff2", C.c_void_p), ("save2", C.c_void_p),
        ("probs", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p),
        ("red2", C.c_void_p), ("red2b", C.c_void_p), ("red1", C.c_void_p),
        ("dzt", C.c_void_p), ("dprobs", C.c_void_p), ("dpool_x", C.c_void_p), ("dx", C.c_void_p), ("dtokp", C.c_void_p),
    ]


_PtrArr = C.c_void_p * MAX_EXPERTS


class XgArgs(C.Structure):
    """avmoe_xg_args: the GEMM-based expert path's row kernels."""
    _fields_ = [
        ("BT", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("Cr", C.c_int32), ("Crp", C.c_int32), ("G", C.c_int32),
        ("E", C.c_int32), ("dtype", C.c_int32), ("bn_mode", C.c_int32), ("_pad", C.c_int32),
        ("ex", Expert * MAX_EXPERTS),
        ("bn2_w", _PtrArr), ("bn2_b", _PtrArr),
        ("x", C.c_void_p), ("xl", C.c_void_p), ("ln1_stats", C.c_void_p),
        ("z", C.c_void_p), ("a", C.c_void_p), ("yd", C.c_void_p), ("ln2_stats", C.c_void_p),
        ("aff1", C.c_void_p), ("save1", C.c_void_p), ("save2", C.c_void_p),
        ("probs", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p),
        ("stats1", C.c_void_p), ("red1", C.c_void_p), ("red2", C.c_void_p),
        ("da", C.c_void_p), ("dz", C.c_void_p), ("dxl", C.c_void_p),
        ("dprobs", C.c_void_p), ("dpool_x", C.c_void_p), ("dx", C.c_void_p), ("res", C.c_void_p),
    ]


class RfBackArgs(C.Structure):
    """avmoe_rf_back_args: the row-fused back pass (bn1 affine .. combine) and its backward."""
    _fields_ = [
        ("BT", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("Cr", C.c_int32), ("Crp", C.c_int32), ("G", C.c_int32),
        ("E", C.c_int32), ("bn_mode", C.c_int32),
        ("ex", Expert * MAX_EXPERTS),
        ("bn2_w", _PtrArr), ("bn2_b", _PtrArr),
        ("z", C.c_void_p), ("aff1", C.c_void_p), ("wu_pack", C.c_void_p), ("ybi", C.c_void_p), ("probs", C.c_void_p),
        ("res", C.c_void_p), ("out", C.c_void_p), ("dout", C.c_void_p), ("yd", C.c_void_p), ("act", C.c_void_p),
        ("dprobs", C.c_void_p), ("red2", C.c_void_p),
    ]


class RfFrontArgs(C.Structure):
    """avmoe_rf_front_args: the row-fused front pass (latent attention of a row .. down projection) and its backward."""
    _fields_ = [
        ("BT", C.c_int32), ("N", C.c_int32), ("C", C.c_int32), ("Cr", C.c_int32), ("Crp", C.c_int32), ("G", C.c_int32),
        ("E", C.c_int32), ("T", C.c_int32), ("has_ln", C.c_int32), ("_pad", C.c_int32),
        ("ex", Expert * MAX_EXPERTS),
        ("tok_s2", _PtrArr), ("tok_xr", _PtrArr),
        ("wd_pack", C.c_void_p), ("wd_csum", C.c_void_p), ("c2", C.c_void_p), ("wdt_pack", C.c_void_p),
        ("x", C.c_void_p), ("z", C.c_void_p), ("dz", C.c_void_p), ("dpool_x", C.c_void_p), ("dx", C.c_void_p), ("xl", C.c_void_p),
        ("a2", _PtrArr), ("ds2", _PtrArr), ("ln1_stats", C.c_void_p), ("dres", C.c_void_p),
    ]


class BnArgs(C.Structure):
    _fields_ = [
        ("E", C.c_int32), ("C", C.c_int32), ("Cr", C.c_int32), ("G", C.c_int32), ("train", C.c_int32), ("_pad", C.c_int32),
        ("n", C.c_int64),
        ("w1", _PtrArr), ("b1", _PtrArr), ("w2", _PtrArr), ("b2", _PtrArr),
        ("rm1", _PtrArr), ("rv1", _PtrArr), ("rm2", _PtrArr), ("rv2", _PtrArr),
        ("wu", _PtrArr),
        ("stats1", C.c_void_p), ("m1", C.c_void_p), ("m2", C.c_void_p),
        ("aff1", C.c_void_p), ("save1", C.c_void_p), ("aff2", C.c_void_p), ("save2", C.c_void_p),
    ]
