"""oracle/render_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT.

CPU restatement of the evaluation renderer with alive-ray rounds (csrc/render.cu + Stage0Trainer.render), i.e. of the reference's
inference loop (NeRFRenderer.render, nerf/renderer.py:749-802) with the product's round schedule instead of the reference's
`n_step = max(min(N // n_alive, 8), 1)`: per round
    plan      n_alive = survivors of the previous round, n_step = min(schedule[round], capacity // n_alive) (>= 1)
    march     raymarching_oracle.march_rays (raymarching.cu:713-828) for the alive rays from their current t
    evaluate  the field on the slab (zero rows behind a ray that ran out)
    composite raymarching_oracle.composite_rays (raymarching.cu:842-924); rays whose slab ended early (zero tail or T < T_thresh) die
and further rounds of the widest slab while rays are left; finally image += (1 - weights_sum) * bg (renderer.py:804).
The operators it is built from are pinned to the reference kernels' golden vectors (tests/test_oracle_golden.py)."""
import numpy as np
import torch

from . import raymarching_oracle as R

SCHEDULE = (8, 8, 16, 16, 32, 64, 128, 256, 512)


def render_rounds(field, rays_o, rays_d, bits, cfg, bg_color, shading="full", amp=True, schedule=SCHEDULE, capacity=None, more=512):
    """-> dict(image [N,3], weights_sum [N], depth [N], rounds, rows) (torch float32 / ints); capacity = sample rows per round (default 16 N)"""
    ro, rd = rays_o.numpy().astype(np.float32), rays_d.numpy().astype(np.float32)
    N = ro.shape[0]
    b = cfg["bound"]
    cap = int(capacity or 16 * N)
    assert cap >= N
    nears, fars = R.near_far_from_aabb(ro, rd, [-b, -b, -b, b, b, b], cfg.get("min_near", 0.05))
    ws = np.zeros(N, np.float32); depth = np.zeros(N, np.float32); image = np.zeros((N, 3), np.float32)
    rays_t = nears.copy()
    alive = np.nonzero(nears < fars)[0].astype(np.int32)
    rounds = rows = 0
    widths = list(schedule)
    while True:
        if not widths:
            if len(alive) == 0:
                break
            widths = [more]
        width = widths.pop(0)
        n_alive = len(alive)
        n_step = width if n_alive * width <= cap else max(1, cap // max(n_alive, 1))
        rounds += 1
        rows += n_alive * n_step
        if n_alive == 0:
            continue                                   # the device runs the (empty) round as well
        xyzs, dirs, ts = R.march_rays(n_alive, n_step, alive, rays_t, ro, rd, b, cfg.get("contract", False), bits.numpy(), cfg["C"], cfg["H"],
                                      nears, fars, np.zeros(n_alive, np.float32), cfg.get("dt_gamma", 0.0), cfg.get("max_steps", 1024))
        x = torch.from_numpy(xyzs); d = torch.from_numpy(dirs)
        d = d / torch.sqrt(torch.clamp((d * d).sum(-1, keepdim=True), min=1e-20))
        with torch.no_grad():
            sig, rgb, _ = field(x, d, shading, amp)
        alive2, rays_t, ws, depth, image = R.composite_rays(n_alive, n_step, alive, rays_t, sig.float().numpy(), rgb.float().numpy(), ts, ws, depth, image,
                                                            cfg.get("T_thresh", 1e-4))
        alive = alive2[alive2 >= 0].astype(np.int32)
    image = torch.from_numpy(image) + (1 - torch.from_numpy(ws)).unsqueeze(-1) * bg_color
    return dict(image=image, weights_sum=torch.from_numpy(ws), depth=torch.from_numpy(depth), rounds=rounds, rows=rows)
