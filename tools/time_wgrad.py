import sys, torch
sys.path.insert(0, '.')
from cunet_b200 import lib
from tests.test_gpu_conv_bwd import make_case, fill_grad_src
from tests.test_gpu_conv_fwd import fill_concat
lib.load()
dtype = lib.BF16
n,h,w,seg_c,ups,cout,taps,mode = 24,64,64,[128,128,32,32],[1,0,0,0],128,1,"bn"
cs = make_case(lib, dtype, n,h,w,seg_c,ups,cout,taps,mode,None)
dw = torch.zeros(cout, cs["cin"], taps, device="cuda")
for nsplit in [0, 148]:
    p = lib.ConvWgradParams()
    fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
    fill_grad_src(p.dy, cs, mode)
    p.N,p.H,p.W,p.taps,p.Cout = n,h,w,taps,cout
    p.dw,p.nsplit,p.dtype = dw.data_ptr(),nsplit,dtype
    for _ in range(3): lib.conv_wgrad(p)
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): lib.conv_wgrad(p)
    e1.record(); torch.cuda.synchronize()
    print("nsplit", nsplit, "us %.1f" % (e0.elapsed_time(e1)*100))
