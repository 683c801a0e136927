"""Per-launch CUDA-event timing of the stage-1 texture step (antialiased and plain), L2 flushed before every launch.
Run on the GPU box: python profiles/s1_stage_probe.py"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from nerf2mesh_b200 import _lib, synthetic as S                      # noqa: E402
from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer        # noqa: E402
from nerf2mesh_b200.stage1 import Stage1Trainer                      # noqa: E402
from nerf2mesh_b200.train_synthetic import full_image_rays           # noqa: E402
R = S                                                                # mesh / projection builders

h0 = w0 = 800
t0 = Stage0Trainer(Stage0Config(bound=1.0, num_rays=1024, max_samples=1024 * 128), seed=0)
v, f = R.icosphere(7)
cam = S.orbit_cameras(8, radius=2.35, seed=3)[0, :3, 3].numpy().astype(np.float64)
pose = torch.from_numpy(S.look_at_pose(cam).astype(np.float32))
intr = S.lego_intrinsics(h0, w0)
_, rd = full_image_rays(pose, intr, h0, w0)
mvp = R.perspective_mvp(cam, fovy=2 * np.arctan(0.5 * h0 / intr[1]), aspect=w0 / h0); mvp[1] *= -1
g = torch.Generator().manual_seed(0)
gt = torch.rand(h0 * w0, 4, generator=g); gt[:, 3] = 1.0
view = (torch.from_numpy(mvp).cuda(), rd.cuda(), gt.cuda(), torch.rand(h0 * w0, 3, generator=g).cuda())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

real_call = _lib.call
res = {}
for aa in (False, True):
    s1 = Stage1Trainer(t0, torch.from_numpy(v), torch.from_numpy(f), h0, w0, ssaa=2, antialias=aa)
    for _ in range(3):
        s1.step(*view)
    torch.cuda.synchronize()
    acc = {}

    def timed(name, *args):
        flush.fill_(0)
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); real_call(name, *args); z.record()
        torch.cuda.synchronize()
        acc.setdefault(name, []).append(a.elapsed_time(z))

    import nerf2mesh_b200.stage1 as m1, nerf2mesh_b200.raster as mr, nerf2mesh_b200.stage0 as m0
    for mod in (m1, mr, m0):
        mod.call = timed
    reps = 5
    for _ in range(reps):
        s1.step(*view)
    for mod in (m1, mr, m0):
        mod.call = real_call
    res["antialias" if aa else "plain"] = {k: round(sum(x) / reps * 1e3, 1) for k, x in acc.items()}          # us per step
    res[("antialias" if aa else "plain") + "_sum_us"] = round(sum(sum(x) for x in acc.values()) / reps * 1e3, 1)
    # the step as the bench runs it (eager, no flush)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        s1.step(*view)
    z.record(); torch.cuda.synchronize()
    res[("antialias" if aa else "plain") + "_step_us"] = round(a.elapsed_time(z) / 20 * 1e3, 1)
    del s1
print(json.dumps(res, indent=1))
