"""The C-ABI library loads on a CPU-only box and exports every symbol include/cunet_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cunet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cunet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from cunet_b200 import lib
    so = ctypes.CDLL(lib.LIB_PATH)
    decl = _declared_symbols()
    assert len(decl) >= 4
    for name in decl:
        assert hasattr(so, name), name
    assert sorted(lib.EXPORTS) == decl
    assert lib.load().cunet_abi_version() >= 1


def test_missing_library_fails_loudly(monkeypatch):
    from cunet_b200 import lib
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/libcunet_b200.so")
    try:
        lib.load()
    except lib.CunetError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("expected CunetError")
