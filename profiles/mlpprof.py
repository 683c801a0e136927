"""Phase profile of the pipelined MLP backward (k_mlp_bwd2): block 0 stamps clock64() at every hand-over of its first
tile (n2m_s0_set_prof).  Prints cycle deltas for tile group 0, the issuer's view of group 0, and tile group 1.
    python profiles/mlpprof.py"""
import sys; sys.path.insert(0, '/root/repo')
import ctypes, torch
import bench
from nerf2mesh_b200 import _lib
from nerf2mesh_b200._lib import P, call, ptr
from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer

_lib.register({"n2m_s0_set_prof": [P]})
tr = Stage0Trainer(Stage0Config(bound=1.0, num_rays=bench.NUM_RAYS, max_samples=bench.NUM_RAYS * 128), seed=0)
host, grid, bits = bench.make_batches(2, 1000, True)
tr.set_occupancy(bits, grid)
b = {k: v.cuda() for k, v in host[0].items()}
for _ in range(3):
    tr.step(b["ro"], b["rd"], b["gt"], b["bg"], b["noises"], use_graph=False)
torch.cuda.synchronize()
buf = torch.zeros(256, dtype=torch.int64, device="cuda")
call("n2m_s0_set_prof", ptr(buf))
tr.mlp_bwd(); torch.cuda.synchronize()
call("n2m_s0_set_prof", ctypes.c_void_p(0))
v = buf.tolist()
def show(name, a):
    a = [x for x in a if x]
    if not a: print(name, "empty"); return
    print(name, "n=%d total=%d" % (len(a), a[-1] - a[0]))
    print("   deltas:", [a[i + 1] - a[i] for i in range(len(a) - 1)])
    return a
g0 = show("group0 (start, tma, [ready, done]*, end)", v[0:32])
iss = show("issuer/g0 ([ready seen, committed]*)", v[32:64])
g1 = show("group1", v[64:96])
if g0 and iss:
    t0 = g0[0]
    print("group0 rel:", [x - t0 for x in g0])
    print("issuer rel:", [x - t0 for x in iss])
    if g1: print("group1 rel:", [x - t0 for x in g1])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for pip in (1, 0):
    call("n2m_s0_set_mlp_bwd_pipelined", pip)
    tr.mlp_bwd(); torch.cuda.synchronize()
    e0.record()
    for _ in range(20): tr.mlp_bwd()
    e1.record(); torch.cuda.synchronize()
    print("mlp_bwd pipelined=%d: %.1f us (warm L2)" % (pip, e0.elapsed_time(e1) * 50))
e0.record()
for _ in range(20): tr.mlp_fwd()
e1.record(); torch.cuda.synchronize()
print("mlp_fwd: %.1f us (warm L2)" % (e0.elapsed_time(e1) * 50))
