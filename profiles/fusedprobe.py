"""Times each role of the fused backward kernel alone (csrc/fused.cu, n2m_s0_set_fused_debug) against the two stand-alone kernels.
    python profiles/fusedprobe.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch
from nerf2mesh_b200 import _lib, synthetic as S
from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer
sys.path.insert(0, '/root/repo')
import bench

tr = Stage0Trainer(Stage0Config(bound=1.0, num_rays=4096, max_samples=4096 * 128), seed=0)
batches, grid, bits = bench.make_batches(1, 1000, False)
tr.set_occupancy(bits, grid)
b = {k: v.cuda() for k, v in batches[0].items()}
tr.rays_o.copy_(b["ro"]); tr.rays_d.copy_(b["rd"]); tr.gt.copy_(b["gt"]); tr.bg.copy_(b["bg"]); tr.noises.copy_(b["noises"])
tr.march(); tr.encode_fwd(); tr.mlp_fwd(); tr.composite_loss()
torch.cuda.synchronize()
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")

def timeit(fn, reps=5):
    t = 0.0
    for _ in range(reps):
        flush.fill_(0.0)
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); z.record(); torch.cuda.synchronize()
        t += a.elapsed_time(z) / reps
    return t * 1e3

print("M", int(tr.counters[1].item()))
print("mlp_bwd            us", round(timeit(tr.mlp_bwd), 1))
print("encode_bwd         us", round(timeit(tr.encode_bwd), 1))
for mode, name in ((0, "fused"), (1, "fused, no REDs"), (2, "fused, no MMA rounds"), (3, "fused, neither")):
    _lib.call("n2m_s0_set_fused_debug", mode)
    print(f"{name:24s} us", round(timeit(tr.bwd_fused), 1))
_lib.call("n2m_s0_set_fused_debug", 0)
