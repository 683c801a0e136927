"""CPU: the antialias oracle (oracle/antialias_oracle.py) against hand-computable cases, invariants of the published algorithm and
finite differences of its own forward pass (nvdiffrast itself cannot be run here: parity unpinned, see the oracle's header)."""
import numpy as np

from oracle import antialias_oracle as A
from oracle import raster_oracle as R


def _quad(x_edge_px, W=8, H=4, z=0.0):
    """a quad covering pixel columns left of the vertical line x = x_edge_px (pixel units), full height; clip space == NDC"""
    xe = x_edge_px / W * 2 - 1
    pos = np.array([[-1, -1, z, 1], [xe, -1, z, 1], [xe, 1, z, 1], [-1, 1, z, 1]], np.float64)
    tri = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return pos, tri


def test_vertical_silhouette_edge_known_blend():
    W, H = 8, 4
    for xe, P_col, Q_col in [(3.8, 3, 4), (4.3, 3, 4)]:
        pos, tri = _quad(xe, W, H)
        rast = R.rasterize(pos, tri, H, W)
        assert (rast[..., 3] > 0).sum() == H * (int(np.floor(xe - 0.5)) + 1)
        color = np.zeros((H, W, 1)); color[rast[..., 3] > 0] = 1.0
        out = A.antialias(color, rast, pos, tri)
        # the boundary edge x = xe lies between the centres of columns 3 (x = 3.5, covered) and 4 (x = 4.5): t = xe - 3.5
        t = xe - 3.5
        exp = color.copy()
        # only the edges of the FOREGROUND PIXEL'S OWN triangle are examined (as in the library): rows whose column-3 pixel shows
        # triangle 1 (which does not own the silhouette edge v1-v2) stay unblended
        own = rast[:, P_col, 3] == 1
        assert own[0] and own.sum() >= H - 1
        if t > 0.5:
            exp[own, Q_col, 0] += (t - 0.5) * 1.0            # foreground bleeds into the empty pixel
        else:
            exp[own, P_col, 0] += (t - 0.5) * 1.0            # the covered pixel loses what the surface does not reach
        assert np.allclose(out, exp, atol=1e-12), (xe, out[:, :, 0])
        # exact coverage of the two straddling pixels: covered pixel spans [3, 4], empty one [4, 5]
        assert np.isclose(out[0, 3, 0] + out[0, 4, 0], np.clip(xe - 3, 0, 1) + np.clip(xe - 4, 0, 1))
        # the diagonal (interior) edge of the quad changes nothing, columns away from the silhouette are untouched
        assert np.array_equal(out[:, :3], color[:, :3]) and np.array_equal(out[:, 5:], color[:, 5:])


def test_interior_edges_are_not_silhouettes_and_folds_are():
    v, f = R.icosphere(2)
    mvp = R.perspective_mvp([1.6, 0.9, 1.1])
    pos = np.concatenate([v, np.ones((len(v), 1), np.float32)], 1) @ mvp.T
    H = W = 48
    rast = R.rasterize(pos, f, H, W)
    hits = A.pairs(rast, pos, f)
    ids = rast[..., 3].reshape(-1)
    # on a closed convex mesh the only silhouette edges seen from outside are on the outline: every blended pair has a background pixel
    assert len(hits) > 40
    assert all(ids[h["P"]] > 0 and ids[h["Q"]] == 0 for h in hits)
    assert all(-0.5 < h["alpha"] < 0.5 for h in hits)
    # uniform colour on both sides of an edge => no change
    out = A.antialias(np.ones((H, W, 3)), rast, pos, f)
    assert np.array_equal(out, np.ones((H, W, 3)))
    # antialiased coverage is closer, pixel by pixel, to the exact (8 x 8 super-sampled) coverage than the binary mask is -- on the
    # pixels the operator touches (it only sees edges of the foreground pixel's own triangle, so it does not find every outline pixel)
    cov = (rast[..., 3:4] > 0).astype(np.float64)
    aa = A.antialias(cov, rast, pos, f)
    fine = (R.rasterize(pos, f, H * 8, W * 8)[..., 3] > 0).reshape(H, 8, W, 8).mean((1, 3))[..., None]
    ch = aa != cov
    assert ch.sum() > 20
    assert np.abs(aa - fine)[ch].mean() < 0.6 * np.abs(cov - fine)[ch].mean(), (np.abs(aa - fine)[ch].mean(), np.abs(cov - fine)[ch].mean())
    assert np.all(aa >= -1e-12) and np.all(aa <= 1 + 1e-12)


def test_backward_matches_finite_differences():
    rng = np.random.default_rng(1)
    v, f = R.icosphere(1)
    # a second, smaller sphere in front: silhouettes over a covered background (both pixels of a pair covered)
    v2 = v * 0.45 + np.array([0.55, 0.3, 0.35], np.float32)
    vv = np.concatenate([v, v2]).astype(np.float64); ff = np.concatenate([f, f + len(v)])
    mvp = R.perspective_mvp([1.6, 0.9, 1.1]).astype(np.float64)
    pos = np.concatenate([vv, np.ones((len(vv), 1))], 1) @ mvp.T
    H, W = 40, 36
    rast = R.rasterize(pos, ff, H, W)
    color = rng.random((H, W, 3))
    gout = rng.standard_normal((H, W, 3))
    hits = A.pairs(rast, pos, ff)
    ids = rast[..., 3].reshape(-1)
    assert sum(ids[h["Q"]] > 0 for h in hits) > 5 and sum(ids[h["Q"]] == 0 for h in hits) > 20
    gc, gp = A.antialias_backward(gout, color, rast, pos, ff, pos_gradient_boost=1.0)
    loss = lambda c, p: (A.antialias(c, rast, p, ff) * gout).sum()            # rast held fixed, as in the operator
    # colours: the operator is linear in them
    dc = rng.standard_normal(color.shape)
    assert np.isclose((gc * dc).sum(), loss(color + dc, pos) - loss(color, pos), rtol=1e-9)
    # positions: central differences on the vertices that received a gradient (t is smooth in x, y, w while no decision flips)
    touched = np.nonzero(np.abs(gp).sum(1) > 0)[0]
    assert len(touched) > 10 and np.all(gp[:, 2] == 0)
    eps = 1e-6
    worst = 0.0
    for vtx in touched[:12]:
        for k in (0, 1, 3):
            p1, p2 = pos.copy(), pos.copy()
            p1[vtx, k] += eps; p2[vtx, k] -= eps
            fd = (loss(color, p1) - loss(color, p2)) / (2 * eps)
            worst = max(worst, abs(fd - gp[vtx, k]) / (abs(gp[vtx]).max() + 1e-12))
    assert worst < 1e-5, worst
    # pos_gradient_boost scales the position gradient only
    gc2, gp2 = A.antialias_backward(gout, color, rast, pos, ff, pos_gradient_boost=3.0)
    assert np.array_equal(gc, gc2) and np.allclose(gp2, 3.0 * gp)
