import sys, torch
sys.path.insert(0, "/root/repo")
import cti_b200
from cti_b200 import kernels as K_
torch.manual_seed(0)
DEV="cuda"
def rel(a,b): return ((a.float()-b.float()).abs().max()/b.float().abs().max()).item()
for (M,N,K) in ((512,512,256),(1000,768,320),(51200,1024,2048)):
    a=torch.randn(M,K,device=DEV).bfloat16(); b=(torch.randn(N,K,device=DEV)/K**0.5).bfloat16(); bias=torch.randn(N,device=DEV)
    ref=a.float()@b.float().t()+bias
    ob,of=K_.gemm(a,b,M,N,K,bias=bias,out_bf16=True,out_f32=True,tile_n=512)
    torch.cuda.synchronize()
    print("fwd",M,N,K,rel(of,ref),rel(ob,ref),flush=True)
    w=(torch.randn(K,N,device=DEV)/K**0.5).bfloat16()
    _,of=K_.gemm(a,w,M,N,K,b_mn=True,out_bf16=False,out_f32=True,tile_n=512)
    torch.cuda.synchronize(); print("dgrad",rel(of,a.float()@w.float()),flush=True)
    dz=torch.randn(M,N,device=DEV).bfloat16(); x=torch.randn(M,K,device=DEV).bfloat16()
    acc=torch.zeros(N,K,device=DEV)
    K_.gemm(dz,x,N,K,M,a_mn=True,b_mn=True,accum_f32=acc,k_splits=3,tile_n=512)
    torch.cuda.synchronize(); print("wgrad",rel(acc,dz.float().t()@x.float()),flush=True)
