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


EPI_STORE, EPI_QUICKGELU, EPI_GELU, EPI_SWIGLU, EPI_F32T, EPI_LMHEAD, EPI_DLOGITS, EPI_F32T_SWIGLU = range(8)

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
    bind(lib)
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().sb_last_error().decode("utf-8", "replace")
        raise SpacerError(f"{what}: {msg}" if what else msg)


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("ldq", C.c_longlong), ("ldk", C.c_longlong), ("ldv", C.c_longlong),
        ("o", C.c_void_p), ("ldo", C.c_longlong),
        ("lse", C.c_void_p), ("meta", C.c_void_p),
        ("T", C.c_int), ("Tk", C.c_int),
        ("n_heads", C.c_int), ("n_kv_heads", C.c_int), ("head_dim", C.c_int),
        ("scale", C.c_float),
        ("d_o", C.c_void_p), ("lddo", C.c_longlong),
        ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("lddq", C.c_longlong),
        ("dk", C.c_void_p), ("dv", C.c_void_p), ("lddk", C.c_longlong), ("lddv", C.c_longlong),
        ("gqa_ws", C.c_void_p), ("tile_ws", C.c_void_p),
    ]


class SampleArgs(C.Structure):
    """sb_sample_args (include/spacer_b200.h)."""
    _fields_ = [
        ("logits", C.c_void_p), ("ld", C.c_longlong), ("R", C.c_int), ("V", C.c_int),
        ("mode", C.c_int), ("top_p", C.c_float), ("top_k", C.c_int), ("temperature", C.c_float),
        ("repetition_penalty", C.c_float),
        ("seen", C.c_void_p), ("seen_ld", C.c_longlong),
        ("seed", C.c_ulonglong), ("seed_dev", C.c_void_p),
        ("step_ptr", C.c_void_p), ("finished", C.c_void_p), ("out_tokens", C.c_void_p), ("out_ids", C.c_void_p),
        ("out_ld", C.c_longlong), ("out_logprob", C.c_void_p),
        ("eos_ids", C.c_int * 4), ("n_eos", C.c_int), ("pad_id", C.c_int), ("suppress_eos", C.c_int),
    ]


_HEADER = Path(__file__).resolve().parent.parent / "include" / "spacer_b200.h"
_CT = {"p": C.c_void_p, "i": C.c_int, "l": C.c_longlong, "f": C.c_float, "u": C.c_ulonglong, "s": C.c_void_p}


def header_signatures() -> dict:
    """Parse include/spacer_b200.h: {function name: type-code string} with p = pointer, i = int,
    l = long long, f = float, u = unsigned long long, s = stream.  The header is the single source of truth
    for the ABI; the ctypes argtypes are derived from it."""
    import re
    src = re.sub(r"/\*.*?\*/", "", _HEADER.read_text(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"\bint\s+(sb_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        codes = ""
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "sb_stream_t" in a:
                    codes += "s"
                elif "*" in a:
                    codes += "p"
                elif "unsigned long long" in a:
                    codes += "u"
                elif "long long" in a:
                    codes += "l"
                elif "float" in a:
                    codes += "f"
                elif "int" in a:
                    codes += "i"
                else:
                    raise SpacerError(f"cannot parse argument '{a}' of {name} in {_HEADER}")
        sigs[name] = codes
    return sigs


def bind(lib):
    for name, sig in header_signatures().items():
        fn = getattr(lib, name)
        fn.argtypes = [_CT[c] for c in sig]
        fn.restype = C.c_int
    return lib
