import sys, torch
sys.path.insert(0, '.')
from cunet_b200 import lib
from tests.test_gpu_conv_bwd import make_case, fill_grad_src
from tests.test_gpu_conv_fwd import fill_concat
lib.load()
dtype = lib.BF16
import os
n = int(os.environ.get("NB","24")); segsel = os.environ.get("SEGS","up")
h,w,cout,taps,mode = 64,64,128,1,"bn"
seg_c,ups = {"up":([128,128,32,32],[1,0,0,0]), "one":([128],[0]), "two":([128,32,32],[0,0,0]), "big2":([128,128],[0,0])}[segsel]
cs = make_case(lib, dtype, n,h,w,seg_c,ups,cout,taps,mode,None)
dev=cs["dev"]; cin=cs["cin"]
nbytes = lib.pack_dgrad_bytes(cin, taps, cs["cout_pad"], dtype)
wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, cin, taps, cs["cout_pad"])
desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
lib.pack_weights(desc_dev.data_ptr(), 1, dtype)
p = lib.ConvDgradParams()
fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"], True)
fill_grad_src(p.dy, cs, mode)
keep=[]
for i,x in enumerate(cs["srcs"]):
    g0=torch.zeros(x.shape, device=dev, dtype=cs["td"]); st=torch.zeros(2*x.shape[1], dtype=torch.float64, device=dev); keep += [g0,st]
    p.gacc[i].G=g0.data_ptr(); p.gacc[i].gstats=st.data_ptr(); p.gacc[i].ld=x.shape[1]; p.gacc[i].accumulate=1
dg=torch.zeros(cin,device=dev); db=torch.zeros(cin,device=dev)
p.N,p.H,p.W,p.taps=n,h,w,taps
p.wpack_dgrad,p.Cout,p.CoutPad=wpack.data_ptr(),cout,cs["cout_pad"]
p.dgamma,p.dbeta,p.dtype=dg.data_ptr(),db.data_ptr(),dtype
import time; t0=time.time(); lib.conv_dgrad(p); torch.cuda.synchronize(); print("first call ok %.2fs"%(time.time()-t0))
for _ in range(3): lib.conv_dgrad(p)
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): lib.conv_dgrad(p)
e1.record(); torch.cuda.synchronize()
print("dgrad us %.1f" % (e0.elapsed_time(e1)*100))
