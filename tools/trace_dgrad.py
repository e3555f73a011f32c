import os, sys, ctypes, torch
sys.path.insert(0, '.')
os.environ.setdefault("NB", "24"); os.environ.setdefault("SEGS", "up")
from cunet_b200 import lib
L = lib.load()
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
exec(open("tools/time_dgrad.py").read().split("import time; t0")[0])   # builds p
lib.conv_dgrad(p); torch.cuda.synchronize()
L.cunet_debug_dgrad_trace(ctypes.c_void_p(buf.data_ptr()))
buf.zero_(); lib.conv_dgrad(p); torch.cuda.synchronize()
L.cunet_debug_dgrad_trace(None)
t = buf.cpu().tolist()
t0 = min(v for v in t if v > 0)
us = lambda v: (v - t0) / 1.9e3 if v > 0 else -1
print("transformer per tile [start, slots ready, dt_free ok, done] (us):")
for tl in range(6): print("  tile#%d" % tl, ["%.1f" % us(t[tl*8+i]) for i in range(4)])
print("epilogue per item [enter, acc_full, slots ok, compute done, barrier passed, store read done] (us):")
for it in range(18): print("  item#%d" % it, ["%.1f" % us(t[64+it*8+i]) for i in range(6)])
print("mma per item [enter, acc_free ok, w_full ok] (us):")
for it in range(18): print("  item#%d" % it, ["%.1f" % us(t[256+it*4+i]) for i in range(3)])
