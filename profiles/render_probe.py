"""Evaluation renderer: device-side alive-ray rounds (csrc/render.cu) vs all-samples chunks of the training kernels, on an 800 x 800 view of
a briefly trained analytic scene.  Run on the GPU box: python profiles/render_probe.py [iters]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from nerf2mesh_b200 import synthetic as S                                   # noqa: E402
from nerf2mesh_b200.sampler import DeviceRaySampler                          # noqa: E402
from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer                # noqa: E402
from nerf2mesh_b200.train_synthetic import full_image_rays, lr_at, psnr     # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = "cuda"
N = 4096
bricks = S.make_bricks()
poses = S.orbit_cameras(100, seed=0)
intr = S.lego_intrinsics()
tr = Stage0Trainer(Stage0Config(bound=1.0, num_rays=N, max_samples=N * 640), seed=0)
tr.density_grid.zero_()
R = 200
bricks_d = tuple(t.to(dev) for t in bricks)
intr_t = intr * (R / 800.0)
imgs = torch.empty(poses.shape[0], R, R, 4, dtype=torch.uint8, device=dev)
for k in range(poses.shape[0]):
    ro, rd = full_image_rays(poses[k], intr_t, R, R)
    imgs[k] = (S.render_bricks(ro.to(dev), rd.to(dev), bricks_d) * 255).round().clamp(0, 255).to(torch.uint8).view(R, R, 4)
sampler = DeviceRaySampler(poses, intr_t, R, R, imgs)
gd = torch.Generator(device=dev).manual_seed(1)
t0 = time.time()
for it in range(iters):
    if it % 16 == 0:
        tr.update_density_grid()
    ro, rd, gt = sampler.sample(N, generator=gd)
    tr.step(ro, rd, gt, torch.rand(N, 3, device=dev, generator=gd), torch.rand(N, device=dev, generator=gd),
            shading="diffuse" if it < iters // 3 else "full", lr=lr_at(it, iters))
torch.cuda.synchronize()
train_s = time.time() - t0

pose = S.orbit_cameras(3, seed=12345)[1]
ro, rd = full_image_rays(pose, intr, 800, 800)
ro, rd = ro.to(dev), rd.to(dev)
gt = S.render_bricks(ro, rd, bricks_d)
gt_rgb = gt[:, :3] * gt[:, 3:] + (1 - gt[:, 3:])
res = {"train_iters": iters, "train_seconds": round(train_s, 2), "rays": int(ro.shape[0])}
for name, kw in (("all_samples", dict(early_stop=False)), ("alive_rounds", dict()), ("alive_rounds_chunk16k", dict(chunk=16384)),
                 ("alive_rounds_chunk256k", dict(chunk=262144))):
    img, ws, dep = tr.render(ro, rd, bg_color=1.0, **kw)                     # warm-up (allocations)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        img, ws, dep = tr.render(ro, rd, bg_color=1.0, **kw)
    z.record(); torch.cuda.synchronize()
    res[name] = {"ms_per_image": round(a.elapsed_time(z) / 3, 3), "psnr_vs_analytic": round(psnr(img.clamp(0, 1), gt_rgb), 3),
                 "covered": round((ws > 0.5).float().mean().item(), 4)}
    if name == "all_samples":
        ref = img.clone()
    else:
        res[name]["max_abs_diff_vs_all_samples"] = float((img - ref).abs().max().item())
        res[name]["rounds_last_chunk"] = tr.render_rounds
# how many sample rows each way evaluates: all samples = the march's M over the image; alive rounds = sum of rows over rounds
M_total = 0
for a0 in range(0, ro.shape[0], N):
    b0 = min(ro.shape[0], a0 + N)
    o = torch.full((N, 3), 1e6, device=dev); d = torch.ones(N, 3, device=dev)
    o[:b0 - a0] = ro[a0:b0]; d[:b0 - a0] = rd[a0:b0]
    tr.slots[tr.cur].load(o, d, torch.zeros(N, 3, device=dev), torch.ones(N, 3, device=dev), torch.zeros(N, device=dev))
    tr.march()
    M_total += int(tr.counters[0].item())
res["samples_marched_all"] = M_total
print(json.dumps(res, indent=1))
