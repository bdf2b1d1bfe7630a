"""ctypes binding of libspacer_b200.so (the C ABI declared in include/spacer_b200.h).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libspacer_b200.so"


class SpacerError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_longlong), ("a_mn", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("b_mn", C.c_int),
        ("D", C.c_void_p), ("ldd", C.c_longlong),
        ("epilogue", C.c_int), ("k_splits", C.c_int), ("bn", C.c_int),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_longlong),
        ("aux", C.c_void_p), ("ldaux", C.c_longlong),
        ("targets", C.c_void_p), ("lse_part", C.c_void_p), ("tgt_logit", C.c_void_p),
        ("lse", C.c_void_p), ("coef", C.c_void_p),
    ]


EPI_STORE, EPI_QUICKGELU, EPI_GELU, EPI_SWIGLU, EPI_F32T, EPI_LMHEAD, EPI_DLOGITS = range(7)

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load the shared library (building is __graft_entry__.build()'s job, never done implicitly)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise SpacerError(
            f"{_LIB_PATH} not found: build it with `python -m spacer_b200.build` "
            "(spacer_b200 has no CPU/PyTorch fallback path)")
    lib = C.CDLL(str(_LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    lib.sb_last_error.restype = C.c_char_p
    lib.sb_abi_version.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sb_last_error().decode("utf-8", "replace")
        raise SpacerError(f"{what}: {msg}" if what else msg)
