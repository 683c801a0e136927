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


def _ground_scene():
    """camera at the origin looking down -z (OpenGL), a ground quad y = -1 reaching from BEHIND the camera (z = +5) to z = -50, plus a
    wall behind the camera that must not appear"""
    near, far, f = 0.1, 100.0, 1.0 / np.tan(0.3)
    proj = np.array([[f, 0, 0, 0], [0, f, 0, 0], [0, 0, (far + near) / (near - far), 2 * far * near / (near - far)], [0, 0, -1, 0]])
    v = np.array([[-30, -1, 5], [30, -1, 5], [30, -1, -50], [-30, -1, -50],            # ground: two vertices behind the camera
                  [-5, -5, 3], [5, -5, 3], [0, 5, 3]], np.float64)                     # wall entirely behind the camera (all w < 0)
    pos = np.concatenate([v, np.ones((len(v), 1))], 1) @ proj.T
    tri = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], np.int32)
    return pos, tri, v, (near, far, f)


def test_triangles_crossing_the_camera_plane_are_clipped_not_dropped():
    pos, tri, v, (near, far, f) = _ground_scene()
    assert (pos[:2, 3] < 0).all() and (pos[2:4, 3] > 0).all() and (pos[4:, 3] < 0).all()
    H = W = 48
    r = R.rasterize(pos, tri, H, W)
    cov = r[..., 3] > 0
    ys = (np.arange(H) + 0.5) / H * 2 - 1
    xs = (np.arange(W) + 0.5) / W * 2 - 1
    A, B = (far + near) / (near - far), 2 * far * near / (near - far)
    for row in range(H):
        for col in range(0, W, 7):
            Y, X = ys[row], xs[col]
            # the pixel's ray (X / f, Y / f, -1) t meets y = -1 at t = -f / Y (Y < 0), eye depth -t
            hit = Y < 0 and (f / -Y) <= 50 and abs(X / f * (f / -Y)) <= 30
            assert cov[row, col] == hit, (row, col)
            if hit:
                t = f / -Y
                assert abs(r[row, col, 2] - (A * -t + B) / t) < 1e-9
                # perspective-correct interpolation of the world position through (u, v) lands on the ray / plane intersection
                i0, i1, i2 = tri[int(r[row, col, 3]) - 1]
                p = r[row, col, 0] * v[i0] + r[row, col, 1] * v[i1] + (1 - r[row, col, 0] - r[row, col, 1]) * v[i2]
                assert np.allclose(p, [X / f * t, -1.0, -t], atol=1e-8)
    assert not (r[..., 3] == 3).any()                                  # the wall behind the camera is never visible
    assert cov.sum() > 0.3 * H * W


def test_near_and_far_planes_clip_per_pixel():
    pos, tri, v, (near, far, f) = _ground_scene()
    # same ground, camera lifted so that the ground passes through the near plane below the view: nothing closer than `near` survives
    v2 = v.copy(); v2[:4, 1] = -0.05                                    # 0.05 below the eye: crosses z_eye = -near at steep rays
    P = np.array([[f, 0, 0, 0], [0, f, 0, 0], [0, 0, (far + near) / (near - far), 2 * far * near / (near - far)], [0, 0, -1, 0]])
    pos2 = np.concatenate([v2, np.ones((len(v2), 1))], 1) @ P.T
    r = R.rasterize(pos2, tri, 64, 64)
    cov = r[..., 3] > 0
    assert cov.any() and (r[cov, 2] >= -1).all() and (r[cov, 2] <= 1).all()
    ys = (np.arange(64) + 0.5) / 64 * 2 - 1
    # rows whose ray reaches the plane before the near distance are empty: t = 0.05 f / -Y < near  <=>  -Y > 0.05 f / near
    steep = -ys > 0.05 * f / near
    assert not cov[steep].any() and cov[(~steep) & (ys < 0)].any()
