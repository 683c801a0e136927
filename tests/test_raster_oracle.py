"""CPU: the rasterizer / interpolation oracle (oracle/raster_oracle.py) against hand-computable cases and invariants of nvdiffrast's
documented output convention (the library itself cannot be run here: parity unpinned, see the oracle's header)."""
import numpy as np

from oracle import raster_oracle as R


def test_single_triangle_known_answers():
    # clip space == NDC (w = 1): right triangle covering the lower-left half of a 4 x 4 target
    pos = np.array([[-1, -1, 0.25, 1], [1, -1, 0.25, 1], [-1, 1, 0.25, 1]], np.float32)
    tri = np.array([[0, 1, 2]], np.int32)
    r = R.rasterize(pos, tri, 4, 4)
    cov = r[..., 3] > 0
    # pixel centres (x+.5, y+.5) with x + y + 1 <= 4 are inside (the diagonal passes through pixel corners, not centres)
    expect = np.array([[(x + y + 1) <= 4 - 1e-9 or (x + y + 1) < 4 for x in range(4)] for y in range(4)])
    expect = np.array([[(x + 0.5) + (y + 0.5) <= 4.0 for x in range(4)] for y in range(4)])
    assert np.array_equal(cov, expect)
    assert np.all(r[cov, 3] == 1) and np.allclose(r[cov, 2], 0.25)
    # barycentrics at pixel (0, 0): centre (0.5, 0.5) px = NDC (-0.75, -0.75): u (vertex 0) = 1 - 0.125 - 0.125
    assert np.allclose(r[0, 0, :2], [0.75, 0.125])
    # row 0 is y_ndc = -1 (OpenGL orientation): vertex 2 (top-left) side is the LAST row
    assert cov[0].sum() == 4 and cov[3].sum() == 1


def test_depth_test_nearest_wins_and_ids_are_one_based():
    quad = lambda z: np.array([[-1, -1, z, 1], [1, -1, z, 1], [1, 1, z, 1], [-1, 1, z, 1]], np.float32)
    pos = np.concatenate([quad(0.5), quad(-0.2), quad(1.5)])      # the last quad is beyond the far plane
    tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7], [8, 9, 10], [8, 10, 11]], np.int32)
    r = R.rasterize(pos, tri, 8, 8)
    assert np.all(r[..., 3] >= 3) and np.all(r[..., 3] <= 4) and np.allclose(r[..., 2], -0.2)


def test_perspective_correct_barycentrics_and_interpolation_roundtrip():
    rng = np.random.default_rng(0)
    v, f = R.icosphere(2)
    mvp = R.perspective_mvp([1.6, 0.9, 1.1])
    pos = np.concatenate([v, np.ones((len(v), 1), np.float32)], 1) @ mvp.T
    r = R.rasterize(pos, f, 64, 64)
    cov = r[..., 3] > 0
    assert 0.1 < cov.mean() < 0.9
    assert np.all(r[cov, 0] >= -1e-12) and np.all(r[cov, 1] >= -1e-12) and np.all(r[cov, 0] + r[cov, 1] <= 1 + 1e-12)
    # interpolating clip-space w and z perspective-correctly must reproduce the depth the rasterizer stored (z/w is screen-linear,
    # z and w are perspective-linear): z_interp / w_interp == rast[..., 2]
    zw = R.interpolate(pos[:, 2:4], r, f)
    assert np.allclose(zw[cov, 0] / zw[cov, 1], r[cov, 2], atol=1e-9)
    # world positions land on the sphere's facets: radius within the chord error of a 2x-subdivided icosphere
    xyz = R.interpolate(v, r, f)
    rad = np.linalg.norm(xyz[cov], axis=1)
    assert rad.max() <= 0.6 + 1e-6 and rad.min() > 0.6 * 0.97
    # only front faces are visible on a closed convex mesh
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    ctr = v[f].mean(1)
    facing = (n * (np.array([1.6, 0.9, 1.1]) - ctr)).sum(1) > 0
    assert facing[r[cov, 3].astype(int) - 1].all()
    # backward of interpolate == transpose of forward
    g = rng.standard_normal(xyz.shape)
    ga = R.interpolate_backward(g, v.shape, r, f)
    a2 = rng.standard_normal(v.shape)
    assert np.isclose((R.interpolate(a2, r, f) * g).sum(), (ga * a2).sum())
