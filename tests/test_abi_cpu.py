"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(sb_\w+)\s*\(", src, flags=re.M):
            names.add(m.group(1))
    return names


def test_library_exports_declared_symbols():
    from spacer_b200 import _lib
    lib = _lib.load()
    declared = _declared()
    assert "sb_gemm" in declared and len(declared) >= 5
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert lib.sb_abi_version() >= 1
    assert isinstance(lib.sb_last_error(), bytes)


def test_no_cpu_fallback_in_product():
    """Product code must not import the oracle."""
    for path in glob.glob(os.path.join(ROOT, "spacer_b200", "**", "*.py"), recursive=True):
        src = open(path).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), path
