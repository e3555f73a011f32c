"""Times the fused 3x3 backward kernel (cunet_conv_bwd3x3) on the bench shape: 128 -> 32, 64x64, batch NB (default 24)."""
import os, sys, torch
sys.path.insert(0, '.')
from cunet_b200 import lib
from tests.test_gpu_conv_bwd import make_case, fill_grad_src
from tests.test_gpu_conv_fwd import fill_concat
lib.load()
dtype = lib.BF16
n = int(os.environ.get("NB", "24")); res = int(os.environ.get("RES", "64"))
cs = make_case(lib, dtype, n, res, res, [128], [0], 32, 9, "bn", None, seed=3)
dev = cs["dev"]
wpack = torch.empty(lib.pack_dgrad_bytes(128, 9, 32, dtype), dtype=torch.uint8, device=dev)
desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), 32, 128, 9, 32)
dd = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev); lib.pack_weights(dd.data_ptr(), 1, dtype)
dp = lib.ConvDgradParams()
fill_concat(dp.inp, cs["srcs"], cs["stats"], cs["counts"], [0], cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
fill_grad_src(dp.dy, cs, "bn")
x = cs["srcs"][0]
g0 = torch.zeros(x.shape, device=dev, dtype=cs["td"]); gst = torch.zeros(256, dtype=torch.float64, device=dev)
dp.gacc[0].G, dp.gacc[0].gstats, dp.gacc[0].ld, dp.gacc[0].accumulate = g0.data_ptr(), gst.data_ptr(), 128, 0
dgamma = torch.zeros(128, device=dev); dbeta = torch.zeros(128, device=dev)
dp.N, dp.H, dp.W, dp.taps = n, res, res, 9
dp.wpack_dgrad, dp.Cout, dp.CoutPad = wpack.data_ptr(), 32, 32
dp.dgamma, dp.dbeta, dp.dtype = dgamma.data_ptr(), dbeta.data_ptr(), dtype
dw = torch.zeros(32, 128, 9, device=dev)
wp = lib.ConvWgradParams()
fill_concat(wp.inp, cs["srcs"], cs["stats"], cs["counts"], [0], cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
fill_grad_src(wp.dy, cs, "bn")
wp.N, wp.H, wp.W, wp.taps, wp.Cout = n, res, res, 9, 32
wp.dw, wp.nsplit, wp.dtype = dw.data_ptr(), 0, dtype
for _ in range(3): lib.conv_bwd3x3(dp, wp)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): lib.conv_bwd3x3(dp, wp)
e1.record(); torch.cuda.synchronize()
rows = n * res * res
print("bwd3x3 N=%d %dx%d: %.1f us/launch (algorithmic bytes %.1f MB)" % (n, res, res, e0.elapsed_time(e1) * 50,
      (rows * (128 * 2 * 2 + 32 * 2 * 2)) / 1e6))
