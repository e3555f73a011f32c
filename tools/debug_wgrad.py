import sys, time, torch
sys.path.insert(0, '.')
from cunet_b200 import lib
from tests import ops_ref
from tests.test_gpu_conv_bwd import make_case, fill_grad_src, reference
from tests.test_gpu_conv_fwd import fill_concat
lib.load()
for dtype_name in ["bf16", "f32"]:
    dtype = lib.BF16 if dtype_name == "bf16" else lib.F32
    for (n,h,w,seg_c,ups,cout,taps,mode) in [(2,16,16,[128],[0],128,1,"plain"), (1,8,8,[128],[0],128,1,"plain")]:
        cs = make_case(lib, dtype, n,h,w,seg_c,ups,cout,taps,mode,None)
        dw = torch.zeros(cout, cs["cin"], taps, device="cuda")
        p = lib.ConvWgradParams()
        fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
        fill_grad_src(p.dy, cs, mode)
        p.N,p.H,p.W,p.taps,p.Cout = n,h,w,taps,cout
        p.dw,p.nsplit,p.dtype = dw.data_ptr(),1,dtype
        torch.cuda.synchronize(); t0=time.time()
        lib.conv_wgrad(p)
        torch.cuda.synchronize(); t1=time.time()
        _,_,_,dw_ref = reference(cs,n,h,w,ups,mode)
        dw2 = dw.reshape(cout, cs["cin"])
        ref = dw_ref.reshape(cout, cs["cin"])
        print(dtype_name, (n,h,w), "time %.3fs"%(t1-t0), "nan frac", torch.isnan(dw2).float().mean().item(),
              "absmax got/ref", dw2[~torch.isnan(dw2)].abs().max().item() if (~torch.isnan(dw2)).any() else None, ref.abs().max().item())
        print(" got[0,:6]", dw2[0,:6].tolist()); print(" ref[0,:6]", ref[0,:6].tolist())
        print(" got[:6,0]", dw2[:6,0].tolist()); print(" ref[:6,0]", ref[:6,0].tolist())
        # check if transposed / permuted relation
        d = (dw2-ref).abs()
        print(" err by 32x32 block (max):")
        print((d.reshape(4,32,4,32).amax(dim=(1,3))).tolist())
