import sys; sys.path.insert(0,'/root/repo')
import torch, ctypes
from nerf2mesh_b200 import _lib
from nerf2mesh_b200._lib import P,U,I,call,ptr,stream
_lib.register({"n2m_tc_bench":[U,U,I,I,U,I,P,P]})
out=torch.zeros(2,dtype=torch.int64,device='cuda')
for (N,ks,a,b) in [(64,4,0,0),(16,4,0,0),(64,4,0,1),(64,8,1,1),(32,8,1,1),(16,8,1,1),(64,1,0,1),(32,1,0,1)]:
    for mode in (0,1):
        call("n2m_tc_bench",N,ks,a,b,200,mode,ptr(out),stream()); torch.cuda.synchronize()
        call("n2m_tc_bench",N,ks,a,b,200,mode,ptr(out),stream()); torch.cuda.synchronize()
        cyc,n=out.tolist()
        print(f"N={N:3d} ksteps={ks} a_mn={a} b_mn={b} mode={mode}: {cyc/n:8.1f} cyc/MMA  {cyc/200:9.1f} cyc/GEMM")
