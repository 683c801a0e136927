"""Thin drivers for the UNMODIFIED reference CUDA kernels in oracle/_ref (raw pybind backends),
following the call protocol of the reference's Python wrappers (raymarching.py / grid.py /
sphere_harmonics.py).  Test infrastructure only."""
import numpy as np
import torch


def near_far(ref, rays_o, rays_d, aabb, min_near):
    N = rays_o.shape[0]
    nears = torch.empty(N, device="cuda"); fars = torch.empty(N, device="cuda")
    ref.near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars)
    return nears, fars


def march_train(ref, rays_o, rays_d, bits, bound, contract, dt_gamma, max_steps, C, H, nears, fars, noises):
    """raymarching.py:229-241: counting pass, .item(), output pass."""
    N = rays_o.shape[0]
    counter = torch.zeros(1, dtype=torch.int32, device="cuda")
    rays = torch.empty(N, 2, dtype=torch.int32, device="cuda")
    ref.march_rays_train(rays_o, rays_d, bits, bound, contract, dt_gamma, max_steps, N, C, H, nears, fars,
                         None, None, None, rays, counter, noises)
    M = int(counter.item())
    xyzs = torch.zeros(M, 3, device="cuda"); dirs = torch.zeros(M, 3, device="cuda"); ts = torch.zeros(M, 2, device="cuda")
    ref.march_rays_train(rays_o, rays_d, bits, bound, contract, dt_gamma, max_steps, N, C, H, nears, fars,
                         xyzs, dirs, ts, rays, counter, noises)
    return xyzs, dirs, ts, rays


def by_ray(x, rays):
    """Reorder per-sample rows into ray order (the reference's offsets follow atomic order)."""
    rays = rays.cpu().numpy()
    x = x.cpu().numpy()
    parts = [x[o:o + c] for o, c in rays]
    return np.concatenate(parts, 0) if parts else x[:0]


def composite_fwd(ref, sigmas, rgbs, ts, rays, T_thresh, alpha_mode):
    M, N = sigmas.shape[0], rays.shape[0]
    weights = torch.zeros(M, device="cuda"); ws = torch.empty(N, device="cuda")
    depth = torch.empty(N, device="cuda"); image = torch.empty(N, 3, device="cuda")
    ref.composite_rays_train_forward(sigmas, rgbs, ts, rays, M, N, T_thresh, alpha_mode, weights, ws, depth, image)
    return weights, ws, depth, image


def composite_bwd(ref, gw, gws, gd, gi, sigmas, rgbs, ts, rays, ws, depth, image, T_thresh, alpha_mode):
    M, N = sigmas.shape[0], rays.shape[0]
    gs = torch.zeros_like(sigmas); gr = torch.zeros_like(rgbs)
    ref.composite_rays_train_backward(gw, gws, gd, gi, sigmas, rgbs, ts, rays, ws, depth, image, M, N, T_thresh,
                                      alpha_mode, gs, gr)
    return gs, gr


def grid_fwd(ref, inputs, emb, offsets, S, H, max_level, gridtype, align, interp, calc_dy_dx):
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = emb.shape[1]
    out = torch.zeros(L, B, C, device="cuda", dtype=emb.dtype)
    dy = torch.zeros(B, L * D * C, device="cuda", dtype=emb.dtype) if calc_dy_dx else None
    ref.grid_encode_forward(inputs, emb, offsets, out, B, D, C, L, max_level, S, H, dy, gridtype, align, interp)
    return out, dy


def grid_bwd(ref, grad, inputs, emb, offsets, S, H, max_level, gridtype, align, interp, dy_dx):
    B, D = inputs.shape
    L = offsets.shape[0] - 1
    C = emb.shape[1]
    gemb = torch.zeros_like(emb)
    ginp = torch.zeros(B, D, device="cuda", dtype=emb.dtype) if dy_dx is not None else None
    ref.grid_encode_backward(grad, inputs, emb, offsets, gemb, B, D, C, L, max_level, S, H, dy_dx, ginp, gridtype, align, interp)
    return gemb, ginp


def sh_fwd(ref, inputs, degree, calc):
    B = inputs.shape[0]
    out = torch.empty(B, degree * degree, device="cuda")
    dy = torch.empty(B, 3 * degree * degree, device="cuda") if calc else None
    ref.sh_encode_forward(inputs, out, B, 3, degree, dy)
    return out, dy
