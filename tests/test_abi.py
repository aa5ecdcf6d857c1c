"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol include/omgsr_b200.h
declares, and the ctypes binding lists the same set. No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "omgsr_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(omgsr_[a-z0-9_]+)\s*\(", txt)))


def _lib():
    from omgsr_b200 import build
    return ctypes.CDLL(build.build())


def test_library_exports_every_declared_symbol():
    lib = _lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/omgsr_b200.h but not exported"


def test_binding_matches_header():
    from omgsr_b200 import capi
    assert sorted(capi.SYMBOLS) == _header_symbols()


def test_version_and_error_string_without_gpu():
    lib = _lib()
    lib.omgsr_last_error.restype = ctypes.c_char_p
    assert lib.omgsr_version() >= 100
    assert isinstance(lib.omgsr_last_error(), bytes)
    # shape validation happens before any CUDA call, so it is checkable on CPU
    rc = lib.omgsr_flash_attn_bf16(None, None, None, None, 1, 1, 128, 128, 96, *([ctypes.c_longlong(8)] * 12),
                                   ctypes.c_float(1.0), None)
    assert rc == -1 and b"head dim" in lib.omgsr_last_error()


def test_product_path_never_imports_oracle():
    """the oracle is test infrastructure: nothing under infer/ or omgsr_b200/ may import it."""
    bad = []
    for top in ("infer", "omgsr_b200"):
        for dp, _, fns in os.walk(os.path.join(ROOT, top)):
            for fn in fns:
                if fn.endswith(".py"):
                    src = open(os.path.join(dp, fn)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
