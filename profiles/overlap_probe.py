"""Can the latency-bound tensor-core MLP kernels share the SMs with the memory-bound gather / scatter kernels?
Times pairs of stage kernels alone and concurrently on two streams (whole 4096-ray batch, data races ignored:
this is a resource-sharing probe, not a correctness run).   python profiles/overlap_probe.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch
import bench
from nerf2mesh_b200._lib import call
from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer

tr = Stage0Trainer(Stage0Config(bound=1.0, num_rays=bench.NUM_RAYS, max_samples=bench.NUM_RAYS * 128), seed=0)
host, grid, bits = bench.make_batches(2, 1000, True)
tr.set_occupancy(bits, grid)
b = {k: v.cuda() for k, v in host[0].items()}
for _ in range(3):
    tr.step(b["ro"], b["rd"], b["gt"], b["bg"], b["noises"], use_graph=False)
torch.cuda.synchronize()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fa, fb=None, reps=10):
    tot = 0.0
    for _ in range(reps):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s1): fa()
        if fb is not None:
            with torch.cuda.stream(s2): fb()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3

stages = {"encode_fwd": tr.encode_fwd, "encode_bwd": tr.encode_bwd, "mlp_fwd": tr.mlp_fwd, "mlp_bwd": tr.mlp_bwd, "adam": tr.adam,
          "tv": tr.tv, "composite": tr.composite_loss}
import nerf2mesh_b200.stage0  # noqa: F401  (registers the hooks)
for carve in (-1, 100, 75):
    pass  # (the carve-out hook was removed in round 2: forcing the max-shared split cost the gather its L1, 114 -> 169 us)
    print("== gather/scatter shared-memory carve-out =", carve)
    alone = {k: timed(f) for k, f in stages.items()}
    print("alone (us, cold L2):", {k: round(v, 1) for k, v in alone.items()})
    for a, c in [("mlp_bwd", "encode_bwd"), ("mlp_bwd", "encode_fwd"), ("mlp_fwd", "encode_fwd"), ("mlp_fwd", "encode_bwd"),
                 ("mlp_bwd", "adam"), ("encode_bwd", "encode_fwd"), ("mlp_bwd", "mlp_fwd"), ("mlp_bwd", "tv")]:
        t = timed(stages[a], stages[c])
        t2 = timed(stages[c], stages[a])
        print(f"{a:10s} || {c:10s}: {t:7.1f} us  (reverse launch order {t2:7.1f})  sum {alone[a] + alone[c]:7.1f}  max {max(alone[a], alone[c]):7.1f}")
