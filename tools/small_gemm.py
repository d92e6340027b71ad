import sys, os
sys.path.insert(0, "/root/repo/tools")
import gemm_sweep as S
for (M,N,K) in ((1024,1024,64),(1024,1024,256),(1024,1024,1024),(1024,1024,4096),(6144,1024,1024),(12288,1024,1024),(12288,512,512)):
    for tn in (128,256):
        S.report("fwd", M,N,K, S.fwd(M,N,K,tile_n=tn), f"tile_n={tn}")
for sp in (1,2,4,8):
    us,_=S.wgrad(1024,1024,1024,splits=sp); S.report("wgrad",1024,1024,1024,us,f"splits={sp}")
S.report("dgrad",1024,1024,1024,S.dgrad(1024,1024,1024,f32=True,aux=False),"prj fp32")
