import sys; sys.path.insert(0,'/root/repo')
import torch, ctypes
from nerf2mesh_b200 import _lib
from nerf2mesh_b200._lib import P,U,I,call,ptr,stream
from profiles.probes import call as probe_call
out=torch.zeros(4,dtype=torch.int64,device='cuda')
for (N,ks,a,b) in [(64,4,0,0),(16,4,0,0),(64,4,0,1),(64,8,1,1),(32,8,1,1),(16,8,1,1),(64,1,0,1),(32,1,0,1)]:
    for mode in (0,1):
        probe_call("n2m_tc_bench",N,ks,a,b,200,mode,ptr(out),stream()); torch.cuda.synchronize()
        probe_call("n2m_tc_bench",N,ks,a,b,200,mode,ptr(out),stream()); torch.cuda.synchronize()
        cyc,n=out.tolist()[:2]
        print(f"N={N:3d} ksteps={ks} a_mn={a} b_mn={b} mode={mode}: {cyc/n:8.1f} cyc/MMA  {cyc/200:9.1f} cyc/GEMM")

# two issuing threads (warps 0 and 1), disjoint accumulators
for (N,ks,a,b) in [(64,4,0,0),(16,4,0,0),(64,8,1,1)]:
    probe_call("n2m_tc_bench",N,ks,a,b,200,2,ptr(out),stream()); torch.cuda.synchronize()
    probe_call("n2m_tc_bench",N,ks,a,b,200,2,ptr(out),stream()); torch.cuda.synchronize()
    c0,n0,c1,n1=out.tolist()
    print(f"2 issuers N={N:3d} ksteps={ks} a_mn={a} b_mn={b}: thread0 {c0/n0:8.1f} cyc/MMA  thread32 {c1/n1:8.1f} cyc/MMA")
