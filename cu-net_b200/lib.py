"""ctypes binding of libcunet_b200.so (the C ABI declared in include/cunet_b200.h).

The product has exactly one compute path: these entry points.  If the shared library is missing or
a call fails, a ``CunetError`` is raised -- there is no CPU / PyTorch fallback.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CUNET_LIB: an alternate build of the same library (kernel experiments); the default is the in-tree build
LIB_PATH = os.environ.get("CUNET_LIB") or os.path.join(_HERE, "libcunet_b200.so")
_lib = None

F32, BF16 = 0, 1
MAX_SEG = 8


class CunetError(RuntimeError):
    pass


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stats", C.c_void_p), ("inv_count", C.c_double),
                ("C", C.c_int), ("ld", C.c_int), ("up", C.c_int), ("reserved", C.c_int)]


class Concat(C.Structure):
    _fields_ = [("seg", Seg * MAX_SEG), ("nseg", C.c_int), ("bn_train", C.c_int),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("rmean", C.c_void_p), ("rvar", C.c_void_p),
                ("eps", C.c_float), ("act_bits", C.c_int)]


class ConvFwdParams(C.Structure):
    _fields_ = [("inp", Concat),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("taps", C.c_int),
                ("wpack", C.c_void_p), ("Cout", C.c_int), ("CoutPad", C.c_int),
                ("out", C.c_void_p), ("out_ld", C.c_int), ("out_fp32", C.c_int),
                ("out_stats", C.c_void_p), ("pool_idx", C.c_void_p), ("pool", C.c_int),
                ("dtype", C.c_int)]


class GradSrc(C.Structure):
    _fields_ = [("g", C.c_void_p), ("t", C.c_void_p), ("stats", C.c_void_p), ("gstats", C.c_void_p),
                ("pool_idx", C.c_void_p), ("inv_count", C.c_double),
                ("C", C.c_int), ("ld", C.c_int), ("mode", C.c_int), ("pooled", C.c_int),
                ("eps", C.c_float), ("reserved", C.c_int)]


class GAcc(C.Structure):
    _fields_ = [("G", C.c_void_p), ("gstats", C.c_void_p), ("ld", C.c_int), ("accumulate", C.c_int)]


class ConvDgradParams(C.Structure):
    _fields_ = [("inp", Concat), ("gacc", GAcc * MAX_SEG), ("dy", GradSrc),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("taps", C.c_int),
                ("wpack_dgrad", C.c_void_p), ("Cout", C.c_int), ("CoutPad", C.c_int),
                ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dtype", C.c_int), ("reserved", C.c_int)]


class ConvWgradParams(C.Structure):
    _fields_ = [("inp", Concat), ("dy", GradSrc),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("taps", C.c_int),
                ("Cout", C.c_int), ("dw", C.c_void_p), ("nsplit", C.c_int), ("dtype", C.c_int),
                ("dw_cin", C.c_int), ("reserved", C.c_int)]


class StemPoolParams(C.Structure):
    _fields_ = [("y", C.c_void_p), ("y_stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("rmean", C.c_void_p), ("rvar", C.c_void_p), ("x", C.c_void_p), ("x_stats", C.c_void_p),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("bn_train", C.c_int), ("eps", C.c_float),
                ("dtype", C.c_int)]


class StemBwdParams(C.Structure):
    _fields_ = [("y", C.c_void_p), ("y_stats", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("dx", GradSrc), ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("dy", C.c_void_p),
                ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("eps", C.c_float), ("dtype", C.c_int),
                ("phase", C.c_int)]


class MseParams(C.Structure):
    _fields_ = [("heads", C.c_void_p * 16), ("dheads", C.c_void_p * 16), ("nheads", C.c_int),
                ("target", C.c_void_p), ("N", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("ld", C.c_int), ("loss", C.c_void_p), ("keys", C.c_void_p), ("grad_scale", C.c_float),
                ("dtype", C.c_int)]


class QuantDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("saved", C.c_void_p), ("grad", C.c_void_p),
                ("Cout", C.c_int), ("Cin", C.c_int), ("taps", C.c_int), ("first_block", C.c_int)]


class PackDesc(C.Structure):
    _fields_ = [("w", C.c_void_p), ("fwd", C.c_void_p), ("dgrad", C.c_void_p),
                ("Cout", C.c_int), ("Cin", C.c_int), ("taps", C.c_int), ("CoutPad", C.c_int)]


def load():
    """Load the shared library (idempotent).  Raises CunetError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CunetError("libcunet_b200.so not found at %s -- run `python __graft_entry__.py` "
                         "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.cunet_last_error.restype = C.c_char_p
    lib.cunet_abi_version.restype = C.c_int
    lib.cunet_pack_fwd_bytes.restype = C.c_long
    lib.cunet_pack_fwd_bytes.argtypes = [C.c_int] * 4
    lib.cunet_pack_dgrad_bytes.restype = C.c_long
    lib.cunet_pack_dgrad_bytes.argtypes = [C.c_int] * 4
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise CunetError("libcunet_b200.so does not export %s (stale build?)" % name)
    _lib = lib
    return lib


# every symbol include/cunet_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "cunet_last_error", "cunet_abi_version",
    "cunet_conv_fwd", "cunet_debug_fwd_v2_min_tiles", "cunet_debug_fwd_v3_min_rows", "cunet_conv_dgrad", "cunet_debug_dgrad_trace", "cunet_conv_wgrad", "cunet_conv_bwd3x3", "cunet_conv_bwd1x1", "cunet_pack_weights", "cunet_pack_fwd_bytes",
    "cunet_pack_dgrad_bytes", "cunet_stem_im2col", "cunet_stem_pool_fwd", "cunet_stem_bwd", "cunet_mse_decode",
    "cunet_decode_finalize", "cunet_bn_running_update", "cunet_rmsprop_step",
    "cunet_quant_forward", "cunet_quant_restore", "cunet_quant_grad", "cunet_quant_input_fwd",
    "cunet_quant_input_bwd",
]


def check(rc, what):
    if rc != 0:
        raise CunetError("%s failed (%d): %s" % (what, rc, load().cunet_last_error().decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def conv_fwd(params):
    check(load().cunet_conv_fwd(C.byref(params), stream_ptr()), "cunet_conv_fwd")


def debug_fwd_v2_min_tiles(min_tiles):
    """Experiment switch of the 1x1 forward dispatch (see include/cunet_b200.h); returns the previous setting."""
    return int(load().cunet_debug_fwd_v2_min_tiles(C.c_int(min_tiles)))


def debug_fwd_v3_min_rows(min_rows):
    """Experiment switch of the third-generation 1x1 forward kernel; returns the previous setting."""
    fn = load().cunet_debug_fwd_v3_min_rows
    fn.restype = C.c_long
    return int(fn(C.c_long(min_rows)))


def conv_dgrad(params):
    check(load().cunet_conv_dgrad(C.byref(params), stream_ptr()), "cunet_conv_dgrad")


def conv_wgrad(params):
    check(load().cunet_conv_wgrad(C.byref(params), stream_ptr()), "cunet_conv_wgrad")


def conv_bwd3x3(dparams, wparams):
    check(load().cunet_conv_bwd3x3(C.byref(dparams), C.byref(wparams), stream_ptr()), "cunet_conv_bwd3x3")


def conv_bwd1x1(dparams, wparams):
    check(load().cunet_conv_bwd1x1(C.byref(dparams), C.byref(wparams), stream_ptr()), "cunet_conv_bwd1x1")


def pack_weights(descs_dev_ptr, ndesc, dtype):
    check(load().cunet_pack_weights(C.c_void_p(descs_dev_ptr), C.c_int(ndesc), C.c_int(dtype), C.c_int(0),
                                    stream_ptr()), "cunet_pack_weights")


def pack_fwd_bytes(cin, taps, cout_pad, dtype):
    return int(load().cunet_pack_fwd_bytes(cin, taps, cout_pad, dtype))


def pack_dgrad_bytes(cin, taps, cout_pad, dtype):
    return int(load().cunet_pack_dgrad_bytes(cin, taps, cout_pad, dtype))
