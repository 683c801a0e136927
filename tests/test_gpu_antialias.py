"""GPU: csrc/antialias.cu (`antialias` with the call surface of nvdiffrast's dr.antialias, reference call sites
nerf/renderer.py:886-887) against the CPU oracle (oracle/antialias_oracle.py, float64 restatement of the published algorithm; parity
with the library itself is unpinned, see its header) on small scenes, plus oracle-free properties at the BASELINE config 5 size."""
import numpy as np
import pytest
import torch

from nerf2mesh_b200 import raster as dr
from oracle import antialias_oracle as A
from oracle import raster_oracle as R

pytestmark = pytest.mark.gpu


def _scene(kind, subdiv):
    v, f = R.icosphere(subdiv)
    if kind == "two":          # a smaller sphere in front of the first: silhouettes over covered pixels (depth decides the foreground)
        v2 = v * 0.45 + np.array([0.55, 0.3, 0.35], np.float32)
        v, f = np.concatenate([v, v2]).astype(np.float32), np.concatenate([f, f + len(v)]).astype(np.int32)
    if kind == "open":         # half of the faces removed: boundary edges (one triangle only)
        f = f[: len(f) // 2]
    return v, f


def _clip(v, cam, aspect=1.0):
    mvp = R.perspective_mvp(cam, aspect=aspect)
    return (np.concatenate([v, np.ones((len(v), 1), np.float32)], 1) @ mvp.T).astype(np.float32)


@pytest.mark.parametrize("kind,subdiv,H,W,cam,C", [("one", 2, 64, 64, [1.6, 0.9, 1.1], 3), ("two", 2, 96, 80, [1.6, 0.9, 1.1], 3),
                                                   ("two", 3, 128, 128, [1.2, 1.3, 0.7], 4), ("open", 2, 72, 72, [0.4, -1.8, 0.6], 1)])
def test_antialias_forward_backward_match_oracle(kind, subdiv, H, W, cam, C):
    rng = np.random.default_rng(3)
    v, f = _scene(kind, subdiv)
    pos = _clip(v, cam, aspect=W / H)
    pos_t = torch.from_numpy(pos).cuda().requires_grad_(True)
    tri_t = torch.from_numpy(f).cuda()
    glctx = dr.RasterizeCudaContext()
    rast, _ = dr.rasterize(glctx, pos_t[None], tri_t, (H, W))
    color = rng.random((H, W, C)).astype(np.float32)
    color_t = torch.from_numpy(color).cuda()[None].requires_grad_(True)
    out = dr.antialias(color_t, rast, pos_t[None], tri_t, pos_gradient_boost=2.0)
    assert out.shape == (1, H, W, C)
    gout = rng.standard_normal((H, W, C)).astype(np.float32)
    (out * torch.from_numpy(gout).cuda()[None]).sum().backward()
    torch.cuda.synchronize()
    # the oracle sees OUR raster (float32 values) and the float32 positions
    rast_np = rast[0].cpu().numpy().astype(np.float64)
    hits = A.pairs(rast_np, pos.astype(np.float64), f)
    assert len(hits) > 30
    if kind == "two":
        ids = rast_np[..., 3].reshape(-1)
        assert sum(ids[h["Q"]] > 0 for h in hits) > 5          # folds over covered pixels are exercised
    ref = A.antialias(color, rast_np, pos.astype(np.float64), f)
    o = out[0].detach().cpu().numpy().astype(np.float64)
    assert np.abs(o - ref).max() <= 2e-4, np.abs(o - ref).max()
    changed = np.abs(ref - color).max(-1) > 0
    assert np.array_equal(np.abs(o - color.astype(np.float64)).max(-1) > 0, changed) or (np.abs(o - ref).max() <= 2e-4 and changed.sum() > 20)
    gc_ref, gp_ref = A.antialias_backward(gout, color, rast_np, pos.astype(np.float64), f, pos_gradient_boost=2.0)
    gc = color_t.grad[0].cpu().numpy().astype(np.float64)
    gp = pos_t.grad.cpu().numpy().astype(np.float64)
    assert np.abs(gc - gc_ref).max() <= 2e-4 * max(1.0, np.abs(gc_ref).max())
    assert np.all(gp[:, 2] == 0)
    assert np.linalg.norm(gp_ref) > 0
    rel = np.linalg.norm(gp - gp_ref) / np.linalg.norm(gp_ref)
    assert rel <= 5e-3, rel


def test_explicit_topology_hash_and_error_paths():
    v, f = _scene("one", 2)
    pos = torch.from_numpy(_clip(v, [1.6, 0.9, 1.1])).cuda()
    tri = torch.from_numpy(f).cuda()
    th = dr.antialias_construct_topology_hash(tri)
    # every edge of a closed manifold mesh has exactly two opposing vertices; the table holds 3F/2 edges
    keys, opp = th.keys.cpu().numpy(), th.opp.cpu().numpy()
    used = keys != -1
    assert used.sum() == 3 * len(f) // 2 and (opp[used] >= 0).all() and (opp[~used] == -1).all()
    edges = {}
    for a, b, c in f:
        for x, y, o in ((a, b, c), (b, c, a), (c, a, b)):
            edges.setdefault((min(x, y), max(x, y)), set()).add(int(o))
    for k, (o0, o1) in zip(keys[used], opp[used]):
        assert edges[(int(k) >> 32, int(k) & 0xffffffff)] == {int(o0), int(o1)}
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos[None], tri, (48, 48))
    c = torch.rand(1, 48, 48, 3, device="cuda")
    a = dr.antialias(c, rast, pos[None], tri, topology_hash=th)
    b = dr.antialias(c, rast, pos[None], tri)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        dr.antialias(c.cpu(), rast, pos[None], tri)
    with pytest.raises(RuntimeError):
        dr.antialias(torch.rand(1, 48, 48, 5, device="cuda"), rast, pos[None], tri)
    with pytest.raises(RuntimeError):
        dr.antialias(torch.rand(1, 40, 48, 3, device="cuda"), rast, pos[None], tri)


@pytest.mark.parametrize("subdiv,min_changed", [(7, 10), (3, 1500)])
def test_full_size_properties(subdiv, min_changed):
    """config 5 size (1600 x 1600; F = 327 680, and 1 280 large faces): uniform colours are untouched, the operator is linear in the
    colours, its colour backward is the transpose of its forward, and on a closed convex mesh only the outline (pixel pairs with one
    background pixel) is touched.  The operator examines the edges of the foreground pixel's OWN triangle only (as published): at
    F = 327 680 the limb triangles are slivers a fraction of a pixel wide and only a few dozen of the ~4 400 outline pixels show the
    triangle that owns the silhouette edge (31 measured); with facets of ~100 px nearly all of them do."""
    v, f = R.icosphere(subdiv)
    pos = torch.from_numpy(_clip(v, np.array([1.7, 0.6, 0.9]))).cuda()
    tri = torch.from_numpy(f).cuda()
    H = W = 1600
    rast, _ = dr.rasterize(dr.RasterizeCudaContext(), pos[None], tri, (H, W))
    th = dr.antialias_construct_topology_hash(tri)
    g = torch.Generator(device="cuda").manual_seed(0)
    ones = torch.ones(1, H, W, 3, device="cuda")
    assert torch.equal(dr.antialias(ones, rast, pos[None], tri, topology_hash=th), ones)
    cov = (rast[..., 3:] > 0).float()
    aa = dr.antialias(cov, rast, pos[None], tri, topology_hash=th)
    assert aa.min().item() >= 0 and aa.max().item() <= 1
    ch = (aa != cov)[0, ..., 0]
    assert min_changed < ch.sum().item() < 20000, ch.sum().item()
    # changed pixels lie on the outline: a 4-neighbour has the other coverage
    c0 = cov[0, ..., 0]
    nb = torch.zeros_like(c0, dtype=torch.bool)
    nb[1:] |= c0[1:] != c0[:-1]; nb[:-1] |= c0[:-1] != c0[1:]; nb[:, 1:] |= c0[:, 1:] != c0[:, :-1]; nb[:, :-1] |= c0[:, :-1] != c0[:, 1:]
    assert (ch & ~nb).sum().item() <= 5
    if subdiv == 3:
        # on the pixels it touches, the antialiased coverage is closer to the exact per-pixel coverage (the same mesh rasterised at 4x
        # the resolution, 16 samples per pixel) than the binary mask is (ratio 0.47 in the CPU oracle's version of this check)
        fine, _ = dr.rasterize(dr.RasterizeCudaContext(), pos[None], tri, (4 * H, 4 * W))
        exact = (fine[0, ..., 3] > 0).float().view(H, 4, W, 4).mean((1, 3))
        e_aa, e_bin = (aa[0, ..., 0] - exact).abs()[ch].mean().item(), (c0 - exact).abs()[ch].mean().item()
        assert e_aa < 0.7 * e_bin, (e_aa, e_bin)
    c1 = torch.rand(1, H, W, 4, device="cuda", generator=g); c2 = torch.rand(1, H, W, 4, device="cuda", generator=g)
    a1 = dr.antialias(c1, rast, pos[None], tri, topology_hash=th); a2 = dr.antialias(c2, rast, pos[None], tri, topology_hash=th)
    a12 = dr.antialias(0.3 * c1 - 1.7 * c2, rast, pos[None], tri, topology_hash=th)
    assert (a12 - (0.3 * a1 - 1.7 * a2)).abs().max().item() <= 1e-5
    assert a1.min().item() >= -1e-6 and a1.max().item() <= 1 + 1e-6
    c1r = c1.clone().requires_grad_(True)
    out = dr.antialias(c1r, rast, pos[None], tri, topology_hash=th)
    gout = torch.randn(1, H, W, 4, device="cuda", generator=g)
    (out * gout).sum().backward()
    lhs = (dr.antialias(c2, rast, pos[None], tri, topology_hash=th).double() * gout.double()).sum().item()
    rhs = (c2.double() * c1r.grad.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs)) + 1e-2, (lhs, rhs)
