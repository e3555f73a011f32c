"""Time (and, with the -DCUNET_TRACE build, trace) the fused 1x1 backward on the bench shapes.

    python tools/time_bwd1x1.py                       # fused vs dgrad + wgrad, several shapes
    CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so python tools/time_bwd1x1.py trace [shape]
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, ".")
from cunet_b200 import lib  # noqa: E402
from tests.test_gpu_conv_bwd import make_case, fill_grad_src  # noqa: E402
from tests.test_gpu_conv_fwd import fill_concat  # noqa: E402

lib.load()
SHAPES = {
    # name: n, h, w, seg_c, ups, cout, dy_mode, cout_pad
    "320up64": (24, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, "bn", None),
    "288up64": (24, 64, 64, [128, 128, 32], [1, 0, 0], 128, "bn", None),
    "256_64": (24, 64, 64, [128, 128], [0, 0], 128, "bn", None),
    "192pool64": (24, 64, 64, [128, 32, 32], [0, 0, 0], 128, "pool", None),
    "192_64": (24, 64, 64, [128, 32, 32], [0, 0, 0], 128, "bn", None),
    "160_64": (24, 64, 64, [128, 32], [0, 0], 128, "bn", None),
    "head68_64": (24, 64, 64, [128], [0], 68, "plain", 80),
    "320up32": (24, 32, 32, [128, 128, 32, 32], [1, 0, 0, 0], 128, "bn", None),
    "192pool32": (24, 32, 32, [128, 32, 32], [0, 0, 0], 128, "pool", None),
    "320up16": (24, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], 128, "bn", None),
    "320up64_b3": (3, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, "bn", None),
}


def make(name):
    n, h, w, seg_c, ups, cout, mode, cout_pad = SHAPES[name]
    dtype = lib.BF16
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, 1, mode, cout_pad)
    dev, cin = cs["dev"], cs["cin"]
    wpack = torch.empty(lib.pack_dgrad_bytes(cin, 1, cs["cout_pad"], dtype), dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, cin, 1, cs["cout_pad"])
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)
    dp, wp = lib.ConvDgradParams(), lib.ConvWgradParams()
    keep = [cs, wpack, desc_dev]
    for q in (dp, wp):
        fill_concat(q.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
        fill_grad_src(q.dy, cs, mode)
    gbytes = 0
    for i, x in enumerate(cs["srcs"]):
        g0 = torch.zeros(x.shape, device=dev, dtype=cs["td"])
        st = torch.zeros(2 * x.shape[1], dtype=torch.float64, device=dev)
        keep += [g0, st]
        dp.gacc[i].G, dp.gacc[i].gstats, dp.gacc[i].ld, dp.gacc[i].accumulate = g0.data_ptr(), st.data_ptr(), x.shape[1], i % 2
        gbytes += x.numel() * 2 * (2 if i % 2 else 1)
    dg, db = torch.zeros(cin, device=dev), torch.zeros(cin, device=dev)
    dw = torch.zeros(cout, cin, device=dev)
    keep += [dg, db, dw]
    dp.N, dp.H, dp.W, dp.taps = n, h, w, 1
    dp.wpack_dgrad, dp.Cout, dp.CoutPad = wpack.data_ptr(), cout, cs["cout_pad"]
    dp.dgamma, dp.dbeta, dp.dtype = dg.data_ptr(), db.data_ptr(), dtype
    wp.N, wp.H, wp.W, wp.taps, wp.Cout = n, h, w, 1, cout
    wp.dw, wp.nsplit, wp.dtype = dw.data_ptr(), 0, dtype
    rows_out = cs["rows"]
    by = sum(x.numel() * 2 for x in cs["srcs"]) + gbytes + rows_out * cs["cout_pad"] * 2 * (1 if mode == "plain" else 2)
    return dp, wp, keep, by


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


def trace(name):
    L = lib.load()
    dp, wp, keep, by = make(name)
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    fn = getattr(L, "cunet_debug_trace_bwd1x1", None)
    if fn is None:
        raise SystemExit("load the trace build through CUNET_LIB (tools/build_trace.sh)")
    lib.conv_bwd1x1(dp, wp)
    torch.cuda.synchronize()
    fn(ctypes.c_void_p(buf.data_ptr()))
    lib.conv_bwd1x1(dp, wp)
    torch.cuda.synchronize()
    fn(None)
    t = buf.cpu().tolist()
    t0 = min(v for v in t if v > 0)
    t = [(v - t0) / 1965.0 if v > 0 else -1.0 for v in t]

    def show(title, base, stride, labels, n=12):
        print("%s [%s] (us)" % (title, ", ".join(labels)))
        for i in range(n):
            vals = t[base + stride * i: base + stride * i + len(labels)]
            if all(v < 0 for v in vals):
                break
            print("  stage#%d %s" % (i, ["%.1f" % v for v in vals]))
    print("== conv_bwd1x1 %s" % name)
    show("producer", 0, 2, ["G/T issued", "x issued"])
    show("transformer", 32, 4, ["start", "G/T landed + dT free", "dT done", "chunk operands done"])
    show("mma", 96, 3, ["dT ready", "first D1 issued", "stage issued"])
    show("epilogue", 144, 4, ["chunk0 acc full", "chunk0 done", "stage done"])
    show("store", 200, 2, ["stores read"])
    print("  tail: stages done %.1f, channel sums added %.1f, wgrad MMAs committed %.1f, D2 complete %.1f, dW added %.1f"
          % (t[228], t[229], t[232], t[230], t[231]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        os.environ.setdefault("CUNET_PDL", "0")
        for nm in (sys.argv[2:] or ["320up64"]):
            trace(nm)
        sys.exit(0)
    names = sys.argv[1:] or list(SHAPES)
    for nm in names:
        dp, wp, keep, by = make(nm)
        tf = timeit(lambda: lib.conv_bwd1x1(dp, wp))

        def sep():
            lib.conv_dgrad(dp)
            lib.conv_wgrad(wp)
        ts = timeit(sep)
        td = timeit(lambda: lib.conv_dgrad(dp))
        print("%-12s fused %7.1f us (%.2f TB/s)   dgrad+wgrad %7.1f us   dgrad alone %7.1f us" % (nm, tf, by / tf / 1e6, ts, td))
