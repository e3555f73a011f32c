import sys, torch
sys.path.insert(0, '.')
from cunet_b200 import lib
from tests.test_gpu_conv_fwd import run_conv_fwd, fill_concat
from tests import ops_ref
lib.load()
dtype = lib.BF16
dev = torch.device("cuda")
def make(n,h,w,seg_c,ups,cout,taps,pool=False):
    """Forward-op parameter struct on random data (returns the struct and the tensors that must stay alive)."""
    g = torch.Generator().manual_seed(0)
    srcs,stats,counts=[],[],[]
    for c,up in zip(seg_c,ups):
        hh,ww=(h//2,w//2) if up else (h,w)
        x=torch.randn(n*hh*ww,c,generator=g).to(dev).to(torch.bfloat16); srcs.append(x); stats.append(ops_ref.tensor_stats(x)); counts.append(n*hh*ww)
    cin=sum(seg_c)
    gamma=torch.rand(cin).to(dev); beta=torch.zeros(cin).to(dev)
    k=3 if taps==9 else 1
    weight=torch.randn(cout,cin,k,k).to(dev)*0.05
    wpack=torch.empty(lib.pack_fwd_bytes(cin,taps,cout,dtype),dtype=torch.uint8,device=dev)
    desc=lib.PackDesc(weight.data_ptr(),wpack.data_ptr(),None,cout,cin,taps,cout)
    dd=torch.frombuffer(bytearray(bytes(desc)),dtype=torch.uint8).to(dev); lib.pack_weights(dd.data_ptr(),1,dtype)
    rows=n*h*w//(4 if pool else 1)
    out=torch.empty(rows,cout,device=dev,dtype=torch.bfloat16); ost=torch.zeros(2*cout,dtype=torch.float64,device=dev)
    pidx=torch.zeros(rows,cout,dtype=torch.uint8,device=dev)
    p=lib.ConvFwdParams(); fill_concat(p.inp,srcs,stats,counts,ups,gamma,beta,gamma,gamma,True)
    p.N,p.H,p.W,p.taps=n,h,w,taps; p.wpack,p.Cout,p.CoutPad=wpack.data_ptr(),cout,cout
    p.out,p.out_ld,p.out_fp32=out.data_ptr(),cout,0; p.out_stats=ost.data_ptr(); p.pool=int(pool); p.pool_idx=pidx.data_ptr(); p.dtype=dtype
    return p, (srcs, stats, gamma, beta, weight, wpack, dd, out, ost, pidx)

def bench(n,h,w,seg_c,ups,cout,taps,pool=False):
    p, keep = make(n,h,w,seg_c,ups,cout,taps,pool)
    cin=sum(seg_c)
    for _ in range(5): lib.conv_fwd(p)
    # eager back-to-back
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): lib.conv_fwd(p)
    e1.record(); torch.cuda.synchronize(); eager=e0.elapsed_time(e1)*5
    gr=torch.cuda.CUDAGraph()
    s=torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        lib.conv_fwd(p)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(gr):
        for _ in range(200): lib.conv_fwd(p)
    gr.replay(); torch.cuda.synchronize(); e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
    print("fwd N=%d %dx%d cin=%d taps=%d: eager %.1f us/launch, graph %.1f us/launch"%(n,h,w,cin,taps,eager,e0.elapsed_time(e1)*5))
import os
if __name__ == "__main__":
    if os.environ.get("CASE","up")=="up":
        bench(24,64,64,[128,128,32,32],[1,0,0,0],128,1)
    else:
        bench(24,64,64,[128],[0],32,9)
