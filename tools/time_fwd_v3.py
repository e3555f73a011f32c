"""Time the 1x1 forward kernels on the bench shapes: third generation (conv_fwd_v3.cu) vs generation 1.

    python tools/time_fwd_v3.py
    CUNET_LIB=$PWD/cu-net_b200/libcunet_b200_trace.so python tools/time_fwd_v3.py trace
"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import tools.time_fwd as tf  # noqa: E402
from cunet_b200 import lib  # noqa: E402

SHAPES = [
    ("320up64", 24, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], False),
    ("288up64", 24, 64, 64, [128, 128, 32], [1, 0, 0], False),
    ("256_64", 24, 64, 64, [128, 128], [0, 0], False),
    ("192_64", 24, 64, 64, [128, 32, 32], [0, 0, 0], False),
    ("192pool64", 24, 64, 64, [128, 32, 32], [0, 0, 0], True),
    ("160_64", 24, 64, 64, [128, 32], [0, 0], False),
    ("320up32", 24, 32, 32, [128, 128, 32, 32], [1, 0, 0, 0], False),
    ("192pool32", 24, 32, 32, [128, 32, 32], [0, 0, 0], True),
    ("320up16", 24, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], False),
    ("192pool16", 24, 16, 16, [128, 32, 32], [0, 0, 0], True),
    ("320up8", 24, 8, 8, [128, 128, 32, 32], [1, 0, 0, 0], False),
    ("320up64_b3", 3, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], False),
]


def timeit(p, reps=50):
    for _ in range(5):
        lib.conv_fwd(p)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        lib.conv_fwd(p)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000.0 / reps


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        L = lib.load()
        name, n, h, w, seg_c, ups, pool = SHAPES[0]
        p, keep = tf.make(n, h, w, seg_c, ups, 128, 1, pool)
        lib.debug_fwd_v3_min_rows(0)
        buf = torch.zeros(512, dtype=torch.int64, device="cuda")
        lib.conv_fwd(p)
        torch.cuda.synchronize()
        L.cunet_debug_trace_fwd_v3(ctypes.c_void_p(buf.data_ptr()))
        lib.conv_fwd(p)
        torch.cuda.synchronize()
        L.cunet_debug_trace_fwd_v3(None)
        t = buf.cpu().tolist()
        t0 = min(v for v in t if v > 0)
        t = [(v - t0) / 1965.0 if v > 0 else -1.0 for v in t]
        print("== conv_fwd_v3", name)
        for title, base, stride, n_ in (("producer x issued", 0, 1, 1), ("transformer start/done", 32, 2, 2),
                                        ("mma acc free/issued", 96, 2, 2), ("epilogue acc full/done", 144, 2, 2),
                                        ("store read", 200, 1, 1)):
            print(title)
            for i in range(12):
                vals = t[base + stride * i: base + stride * i + n_]
                if all(v < 0 for v in vals):
                    break
                print("  stage#%d %s" % (i, ["%.1f" % v for v in vals]))
        print("prologue: kernel start %.2f, after griddep_wait %.2f, coefficients ready %.2f" % (t[296], t[297], t[298]))
        print("transformer stage 4 detail: x landed %.2f | per chunk [start, a_free ok, stores done, fenced+arrived]" % t[299])
        for c in range(3):
            print("  chunk %d %s" % (c, ["%.2f" % v for v in t[300 + 4 * c: 304 + 4 * c]]))
        sys.exit(0)
    want = [a for a in sys.argv[1:] if a != "trace"]
    for name, n, h, w, seg_c, ups, pool in SHAPES:
        if want and name not in want:
            continue
        p, keep = tf.make(n, h, w, seg_c, ups, 128, 1, pool)
        cin = sum(seg_c)
        by = sum(x.numel() * 2 for x in keep[0]) + keep[7].numel() * 2
        lib.debug_fwd_v3_min_rows(0)
        t3 = timeit(p)
        lib.debug_fwd_v3_min_rows(-1)
        t1 = timeit(p)
        lib.debug_fwd_v3_min_rows(0)
        print("%-12s v3 %7.1f us (%.2f TB/s)   gen-1 %7.1f us" % (name, t3, by / t3 / 1e6, t1))
