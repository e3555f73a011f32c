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


def test_argument_validation_happens_on_the_host():
    """Bad arguments are rejected before any CUDA call (so this runs without a GPU) with a message in
    cunet_last_error(); nothing is launched."""
    from cunet_b200 import lib
    so = lib.load()
    so.cunet_last_error.restype = ctypes.c_char_p
    assert so.cunet_conv_bwd3x3(None, None, None) < 0
    assert b"null params" in so.cunet_last_error()
    p = lib.ConvFwdParams()
    p.inp.nseg = 0
    assert so.cunet_conv_fwd(ctypes.byref(p), None) < 0 and b"nseg" in so.cunet_last_error()
    p.inp.nseg, p.taps = 1, 5
    assert so.cunet_conv_fwd(ctypes.byref(p), None) < 0 and b"taps" in so.cunet_last_error()
    d = lib.ConvDgradParams()
    d.inp.nseg, d.taps, d.inp.bn_train = 1, 1, 0
    assert so.cunet_conv_dgrad(ctypes.byref(d), None) < 0 and b"train-mode" in so.cunet_last_error()
    # the experiment switch of the 1x1 forward dispatch round-trips its value
    old = lib.debug_fwd_v2_min_tiles(7)
    assert lib.debug_fwd_v2_min_tiles(old) == 7
