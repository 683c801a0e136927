"""GPU: csrc/raster.cu (rasterize / interpolate, C ABI include/n2m_b200_raster.h) against the CPU oracle (oracle/raster_oracle.py,
float64) on icosphere meshes from several viewpoints, plus size-independent properties at the BASELINE config 5 size
(F ~ 3e5 triangles, 1600 x 1600)."""
import numpy as np
import pytest
import torch

from nerf2mesh_b200 import raster as dr
from oracle import raster_oracle as R

pytestmark = pytest.mark.gpu


def _clip(v, cam, aspect=1.0):
    mvp = R.perspective_mvp(cam, aspect=aspect)
    return (np.concatenate([v, np.ones((len(v), 1), np.float32)], 1) @ mvp.T).astype(np.float32)


@pytest.mark.parametrize("subdiv,H,W,cam", [(2, 96, 96, [1.6, 0.9, 1.1]), (3, 200, 160, [0.2, -1.9, 0.4]), (4, 256, 256, [-1.2, 1.3, -0.8]),
                                             (1, 64, 64, [0.9, 0.1, 0.2])])
def test_rasterize_and_interpolate_match_oracle(subdiv, H, W, cam):
    v, f = R.icosphere(subdiv)
    pos = _clip(v, cam, aspect=W / H)
    ref = R.rasterize(pos, f, H, W)
    glctx = dr.RasterizeCudaContext()
    rast, _ = dr.rasterize(glctx, torch.from_numpy(pos).cuda()[None], torch.from_numpy(f).cuda(), (H, W))
    assert rast.shape == (1, H, W, 4)
    out = rast[0].cpu().numpy().astype(np.float64)
    # ids: identical except where a pixel centre is within fp32 rounding of an edge
    mism = out[..., 3] != ref[..., 3]
    edge = R.edge_distance(pos.astype(np.float64), f, ref)
    edge_o = R.edge_distance(pos.astype(np.float64), f, out)
    assert (mism & (np.minimum(edge, edge_o) > 1e-4)).sum() == 0, int(mism.sum())
    assert mism.mean() < 2e-3
    ok = ~mism & (ref[..., 3] > 0)
    assert ok.mean() > 0.02
    assert np.abs(out[ok, :3] - ref[ok, :3]).max() <= 2e-4, np.abs(out[ok, :3] - ref[ok, :3]).max()
    assert np.all(out[ref[..., 3] == 0][~mism[ref[..., 3] == 0]] == 0)
    # interpolate (forward) on OUR raster vs the oracle on our raster, 3 and 1 attributes (renderer.py:862-863)
    tri_t = torch.from_numpy(f).cuda()
    xyz, _ = dr.interpolate(torch.from_numpy(v).cuda()[None], rast, tri_t)
    assert np.abs(xyz[0].cpu().numpy() - R.interpolate(v, out, f)).max() <= 1e-5
    ones, _ = dr.interpolate(torch.ones(len(v), 1, device="cuda")[None], rast, tri_t)
    assert torch.equal(ones[0, ..., 0] > 0, rast[0, ..., 3] > 0)
    # backward w.r.t. attr
    attr = torch.from_numpy(v).cuda().requires_grad_(True)
    o, _ = dr.interpolate(attr[None], rast, tri_t)
    g = torch.randn_like(o)
    (o * g).sum().backward()
    gref = R.interpolate_backward(g[0].cpu().numpy().astype(np.float64), v.shape, out, f)
    assert np.abs(attr.grad.cpu().numpy() - gref).max() <= 1e-4 * max(1.0, np.abs(gref).max())


def test_full_size_properties():
    """config 5 size: icosphere with 327 680 faces on a 1600 x 1600 target (ssaa 2 of an 800^2 view).  Properties that need no
    oracle: every covered pixel's triangle contains the pixel centre, depth equals the perspective interpolation of z / w, only
    front faces win, coverage equals the analytic silhouette of the sphere, large (queued) triangles agree with small ones."""
    v, f = R.icosphere(7)
    assert f.shape[0] == 327680
    cam = np.array([1.7, 0.6, 0.9])
    pos = _clip(v, cam)
    H = W = 1600
    glctx = dr.RasterizeCudaContext()
    pos_t, tri_t = torch.from_numpy(pos).cuda(), torch.from_numpy(f).cuda()
    rast, _ = dr.rasterize(glctx, pos_t[None], tri_t, (H, W))
    r = rast[0]
    cov = r[..., 3] > 0
    frac = cov.float().mean().item()
    # a sphere of radius rho seen from distance d with vertical fov: silhouette is a disc of angular radius asin(rho / d)
    d = np.linalg.norm(cam); ang = np.arcsin(0.6 / d)
    expect = np.pi * (np.tan(ang) / np.tan(0.6911 / 2)) ** 2 / 4
    assert abs(frac - expect) < 0.01 * expect + 1e-3, (frac, expect)
    ids = (r[..., 3][cov] - 1).long()
    u, vv = r[..., 0][cov], r[..., 1][cov]
    assert (u >= -1e-5).all() and (vv >= -1e-5).all() and (u + vv <= 1 + 1e-5).all()
    zw, _ = dr.interpolate(pos_t[:, 2:4][None].contiguous(), rast, tri_t)
    z = (zw[0, ..., 0] / zw[0, ..., 1])[cov]
    assert (z - r[..., 2][cov]).abs().max().item() <= 2e-5
    vt = torch.from_numpy(v).cuda()
    n = torch.cross(vt[tri_t[:, 1].long()] - vt[tri_t[:, 0].long()], vt[tri_t[:, 2].long()] - vt[tri_t[:, 0].long()], dim=-1)
    ctr = vt[tri_t.long()].mean(1)
    facing = (n * (torch.tensor(cam, device="cuda", dtype=torch.float32) - ctr)).sum(-1) > 0
    assert facing[ids].all()
    # one huge triangle in front of everything goes through the queued (block-per-triangle) path and must win everywhere it covers
    big = np.array([[-0.9, -0.9, -0.5, 1], [0.9, -0.9, -0.5, 1], [0.0, 0.9, -0.5, 1]], np.float32)
    pos2 = torch.cat([pos_t, torch.from_numpy(big).cuda()])
    tri2 = torch.cat([tri_t, torch.tensor([[len(pos), len(pos) + 1, len(pos) + 2]], dtype=torch.int32, device="cuda")])
    rast2, _ = dr.rasterize(glctx, pos2[None], tri2, (H, W))
    won = rast2[0, ..., 3] == f.shape[0] + 1
    ys, xs = torch.meshgrid(torch.arange(H, device="cuda") + 0.5, torch.arange(W, device="cuda") + 0.5, indexing="ij")
    xn, yn = xs / W * 2 - 1, ys / H * 2 - 1
    inside = (yn > -0.9 + 1e-4) & (yn < 0.9 - 2 * (xn.abs()) * 1.0 - 1e-4 + 0.0)          # |x| <= (0.9 - y) / 2
    assert won[inside].all()
    assert torch.equal(rast2[0][~won], rast[0][~won])


def test_compact_covered_matches_boolean_mask():
    v, f = R.icosphere(3)
    pos = _clip(v, [1.6, 0.9, 1.1])
    glctx = dr.RasterizeCudaContext()
    tri_t = torch.from_numpy(f).cuda()
    rast, _ = dr.rasterize(glctx, torch.from_numpy(pos).cuda()[None], tri_t, (128, 128))
    xyz, _ = dr.interpolate(torch.from_numpy(v).cuda()[None], rast, tri_t)
    dirs = torch.randn(128 * 128, 3, device="cuda")
    cnt, pix, pts, pd = dr.compact_covered(rast, xyz.view(-1, 3), dirs)
    mask = (rast[0, ..., 3] > 0).view(-1)
    n = int(cnt.item())
    assert n == int(mask.sum().item())
    order = torch.argsort(pix[:n])
    assert torch.equal(pix[:n][order].long(), torch.nonzero(mask).view(-1))
    assert torch.equal(pts[:n][order], xyz.view(-1, 3)[mask]) and torch.equal(pd[:n][order], dirs[mask])


@pytest.mark.parametrize("H,W", [(64, 64), (200, 320)])
def test_triangles_crossing_the_camera_plane_match_oracle(H, W):
    """a camera INSIDE the mesh (the outer shells of a bound-16 scene, a ground plane under the camera): triangles with one or two
    vertices behind the camera plane are rasterised in homogeneous coordinates -- the part in front of the near plane, with the same
    depth and perspective-correct barycentrics the oracle (tests/test_raster_oracle.py: analytic ground-plane check) produces"""
    rng = np.random.default_rng(0)
    near, far, f = 0.1, 100.0, 1.0 / np.tan(0.3)
    P = np.array([[f * H / W, 0, 0, 0], [0, f, 0, 0], [0, 0, (far + near) / (near - far), 2 * far * near / (near - far)], [0, 0, -1, 0]])
    # ground grid 12 x 12 from z = +6 (behind) to z = -40, a big icosphere around the camera (radius 8, seen from inside), a small one in front
    gx, gz = np.meshgrid(np.linspace(-20, 20, 13), np.linspace(6, -40, 13), indexing="ij")
    gv = np.stack([gx.reshape(-1), np.full(gx.size, -0.2), gz.reshape(-1)], 1)          # 0.2 below the eye: the cells that cross z = 0 are in view
    gf = []
    for i in range(12):
        for j in range(12):
            a = i * 13 + j
            gf += [[a, a + 13, a + 14], [a, a + 14, a + 1]]
    sv, sf = R.icosphere(2, radius=8.0)
    sv = sv + np.array([0.3, -0.2, 0.5], np.float32)
    tv, tf = R.icosphere(2, radius=0.7)
    tv = tv + np.array([0.5, 0.2, -4.0], np.float32)
    v = np.concatenate([gv, sv, tv]).astype(np.float32)
    fcs = np.concatenate([np.array(gf), sf + len(gv), tf + len(gv) + len(sv)]).astype(np.int32)
    pos = (np.concatenate([v, np.ones((len(v), 1), np.float32)], 1) @ P.T).astype(np.float32)
    w = pos[:, 3]
    crossing = ((w[fcs] <= 0).any(1) & (w[fcs] > 0).any(1)).sum()
    assert crossing > 30 and ((w[fcs] <= 0).all(1)).sum() > 30                       # both kinds are present
    ref = R.rasterize(pos, fcs, H, W)
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), torch.from_numpy(pos).cuda()[None], torch.from_numpy(fcs).cuda(), (H, W))
    out = rast[0].cpu().numpy().astype(np.float64)
    assert (ref[..., 3] > 0).mean() > 0.95                                           # the shell around the camera fills the view
    crossing_ids = np.nonzero((w[fcs] <= 0).any(1))[0] + 1
    assert np.isin(ref[..., 3], crossing_ids).mean() > 0.05                          # and crossing triangles are really visible
    mism = out[..., 3] != ref[..., 3]
    # ids may differ only where a pixel centre is within fp32 rounding of an edge (smallest barycentric ~ 0) or two depths tie
    near_edge = np.minimum(np.minimum(ref[..., 0], ref[..., 1]), 1 - ref[..., 0] - ref[..., 1]) < 2e-4
    near_edge_o = np.minimum(np.minimum(out[..., 0], out[..., 1]), 1 - out[..., 0] - out[..., 1]) < 2e-4
    assert (mism & ~(near_edge | near_edge_o)).sum() == 0, int((mism & ~(near_edge | near_edge_o)).sum())
    assert mism.mean() < 5e-3
    ok = ~mism & (ref[..., 3] > 0)
    assert np.abs(out[ok, :2] - ref[ok, :2]).max() <= 5e-4 and np.abs(out[ok, 2] - ref[ok, 2]).max() <= 2e-5
    # world positions through interpolate: the ground pixels lie on y = -0.2
    xyz, _ = dr.interpolate(torch.from_numpy(v).cuda()[None], rast, torch.from_numpy(fcs).cuda())
    ground = torch.from_numpy((out[..., 3] > 0) & (out[..., 3] <= len(gf))).cuda()
    assert ground.any() and (xyz[0][ground][:, 1] + 0.2).abs().max().item() <= 1e-3
