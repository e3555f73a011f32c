"""In-kernel timelines of the second-generation kernels on the bench shapes (CTA 0's clock64() marks).

    bash tools/build_trace.sh
    CUNET_PDL=0 CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so python tools/trace_kernels.py [bwd3x3 fwd3x3 wgrad fwd_v2]

Needs the -DCUNET_TRACE build (the default library compiles the marks to nothing and lacks the setters).  Slot layouts
are documented next to CUNET_TRACE_DECL in each kernel's source.  (conv_dgrad_v2 has its own tool: trace_dgrad.py.)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
os.environ.setdefault("CUNET_PDL", "0")
from cunet_b200 import lib  # noqa: E402

L = lib.load()
CLK_GHZ = 1.965


def capture(setter, launch):
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    fn = getattr(L, setter, None)
    if fn is None:
        raise SystemExit("%s missing: load the trace build through CUNET_LIB (tools/build_trace.sh)" % setter)
    launch()
    torch.cuda.synchronize()
    fn(ctypes.c_void_p(buf.data_ptr()))
    launch()
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().tolist()
    t0 = min(v for v in t if v > 0)
    return [(v - t0) / (CLK_GHZ * 1e3) if v > 0 else -1.0 for v in t]


def show(title, t, rows, base, stride, labels, n=12):
    print("%s  [%s] (us)" % (title, ", ".join(labels)))
    for i in range(n):
        vals = t[base + stride * i: base + stride * i + len(labels)]
        if all(v < 0 for v in vals):
            break
        print("  %s#%d %s" % (rows, i, ["%.1f" % v for v in vals]))


def trace_bwd3x3():
    src = open("tools/time_bwd3x3.py").read().split("for _ in range(3): lib.conv_bwd3x3")[0]
    ns = {}
    exec(src, ns)
    t = capture("cunet_debug_trace_bwd3x3", lambda: lib.conv_bwd3x3(ns["dp"], ns["wp"]))
    print("== conv_bwd3x3, 128 -> 32, 64x64, batch %d" % ns["n"])
    show("producer", t, "stage", 0, 2, ["G/T issued", "x issued"])
    show("transformer", t, "stage", 32, 4, ["start", "landed", "operands free", "done"])
    show("mma", t, "stage", 96, 3, ["operands ready", "D1 free", "issued"])
    show("epilogue", t, "stage", 144, 2, ["D1 full", "done"])
    show("store", t, "stage", 176, 2, ["G ready", "store read done"])
    print("  dW staging starts %.1f, staged %.1f, added to the gradient %.1f" % (t[208], t[209], t[210]))


def trace_fwd(kind):
    import tools.time_fwd as tf
    if kind == "fwd3x3":
        p, keep = tf.make(24, 64, 64, [128], [0], 32, 9)
        setter = "cunet_debug_trace_fwd3x3"
    else:
        lib.debug_fwd_v2_min_tiles(1)
        p, keep = tf.make(24, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1)
        setter = "cunet_debug_trace_fwd_v2"
    t = capture(setter, lambda: lib.conv_fwd(p))
    print("== %s, 64x64, batch 24" % kind)
    show("producer", t, "tile", 0, 1, ["landing issued"])
    show("transformer", t, "tile", 16, 3, ["start", "landed / operand free", "done"])
    show("mma", t, "tile", 64, 2, ["ready", "issued"])
    if kind == "fwd3x3":
        show("epilogue", t, "tile", 96, 3, ["acc full", "TMEM drained", "done"])
    else:
        show("epilogue", t, "tile", 96, 2, ["acc full", "done"])


def trace_wgrad():
    src = open("tools/time_wgrad.py").read().split("for nsplit in")[0]
    ns = {}
    exec(src, ns)
    cs, lb = ns["cs"], ns["lib"]
    from tests.test_gpu_conv_bwd import fill_grad_src
    from tests.test_gpu_conv_fwd import fill_concat
    p = lb.ConvWgradParams()
    fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ns["ups"], cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
    fill_grad_src(p.dy, cs, ns["mode"])
    p.N, p.H, p.W, p.taps, p.Cout = ns["n"], ns["h"], ns["w"], ns["taps"], ns["cout"]
    p.dw, p.nsplit, p.dtype = ns["dw"].data_ptr(), 0, ns["dtype"]
    t = capture("cunet_debug_trace_wgrad_v2", lambda: lb.conv_wgrad(p))
    print("== conv_wgrad_v2, 320 -> 128, 64x64, batch 24")
    show("producer", t, "stage", 0, 1, ["first landing issued"])
    show("transformer", t, "stage", 16, 4, ["start", "dT landed + B free", "dT operand done", "chunk operands done"])
    show("mma", t, "stage", 80, 2, ["B ready", "issued"])
    print("  epilogue starts %.1f, done %.1f" % (t[120], t[121]))


if __name__ == "__main__":
    which = sys.argv[1:] or ["bwd3x3", "fwd3x3", "wgrad", "fwd_v2"]
    for k in which:
        {"bwd3x3": trace_bwd3x3, "wgrad": trace_wgrad, "fwd3x3": lambda: trace_fwd("fwd3x3"),
         "fwd_v2": lambda: trace_fwd("fwd_v2")}[k]()
