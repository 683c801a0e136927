"""GPU: csrc/mcubes.cu (`nerf2mesh_b200.mesh.marching_cubes`, the device-side replacement of the `mcubes.marching_cubes` call of
NeRFRenderer.export_stage0, nerf/renderer.py:526-529) against the CPU oracle (oracle/mcubes_oracle.py; PyMCubes itself cannot be run
here: parity unpinned, see its header), full-size properties at the reference's default resolution 512^3, and the export up to
`mesh_0.ply`."""
import numpy as np
import pytest
import torch

from nerf2mesh_b200 import mesh as M
from oracle import mcubes_oracle as O

pytestmark = pytest.mark.gpu


def _grid(R, dev="cpu"):
    g = torch.linspace(-1, 1, R, device=dev, dtype=torch.float64)
    return torch.meshgrid(g, g, g, indexing="ij")


@pytest.mark.parametrize("kind,R", [("sphere", 33), ("torus", 40), ("noise", 17), ("slab", (9, 20, 31))])
def test_matches_oracle(kind, R):
    if kind == "slab":                                    # anisotropic volume, surface cut by the volume faces
        X, Y, Z = R
        gx, gy, gz = np.meshgrid(np.linspace(-1, 1, X), np.linspace(-1, 1, Y), np.linspace(-1, 1, Z), indexing="ij")
        vol, iso = (np.exp(-3 * (gx * gx + 0.5 * gy * gy)) * 20 + 2 * np.sin(5 * gz)).astype(np.float32), 10.0
    else:
        x, y, z = (t.numpy() for t in _grid(R))
        if kind == "sphere":
            vol, iso = (0.6 - np.sqrt(x * x + y * y + z * z)).astype(np.float32), 0.0
        elif kind == "torus":
            q = np.sqrt(x * x + y * y) - 0.55
            vol, iso = (0.22 - np.sqrt(q * q + z * z)).astype(np.float32), 0.0
        else:                                             # every case of the table, ambiguous faces included
            rng = np.random.default_rng(1)
            vol = np.zeros((R, R, R), np.float32)
            vol[1:-1, 1:-1, 1:-1] = (rng.random((R - 2,) * 3) + (rng.random((R - 2,) * 3) > 0.5)).astype(np.float32)
            iso = 1.0
    v_ref, f_ref = O.marching_cubes(vol.astype(np.float64), iso)
    v, f = M.marching_cubes(torch.from_numpy(vol).cuda(), iso)
    assert v.dtype == torch.float32 and f.dtype == torch.int32
    assert v.shape == v_ref.shape and f.shape == f_ref.shape and f.shape[0] > 50
    assert np.array_equal(f.cpu().numpy().astype(np.int64), f_ref)                     # same triangles in the same order
    assert np.abs(v.cpu().numpy().astype(np.float64) - v_ref).max() <= 1e-5
    if kind != "slab":
        pr = O.mesh_properties(v.cpu().numpy().astype(np.float64), f.cpu().numpy().astype(np.int64))
        assert pr["closed"] and pr["consistent"]


def test_empty_and_error_paths():
    v, f = M.marching_cubes(torch.zeros(8, 8, 8, device="cuda"), 0.5)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    v, f = M.marching_cubes(torch.ones(8, 8, 8, device="cuda"), 0.5)                  # everything inside: no surface either
    assert v.shape == (0, 3) and f.shape == (0, 3)
    with pytest.raises(RuntimeError):
        M.marching_cubes(torch.zeros(8, 8, 8), 0.5)
    with pytest.raises(RuntimeError):
        M.marching_cubes(torch.zeros(8, 8, device="cuda"), 0.5)


def test_full_size_sphere_properties(tmp_path):
    """512^3 (the reference's --mcubes_reso default is 512): closed, consistently oriented, genus 0, analytic area / volume, PLY round trip"""
    R = 512
    x, y, z = _grid(R, "cuda")
    vol = (0.6 - torch.sqrt(x * x + y * y + z * z)).float()
    del x, y, z
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    M.marching_cubes(vol, 0.0)
    e0.record(); v, f = M.marching_cubes(vol, 0.0); e1.record()
    torch.cuda.synchronize()
    print(f"marching cubes 512^3: {e0.elapsed_time(e1):.2f} ms, {v.shape[0]} vertices, {f.shape[0]} triangles")
    h = 2.0 / (R - 1)
    fl = f.long()
    # closedness + orientation: every directed edge once, its reverse once
    e = torch.cat([fl[:, [0, 1]], fl[:, [1, 2]], fl[:, [2, 0]]])
    key = e[:, 0] * v.shape[0] + e[:, 1]
    rkey = e[:, 1] * v.shape[0] + e[:, 0]
    uk, cnt = torch.unique(key, return_counts=True)
    assert (cnt == 1).all() and torch.equal(uk, torch.unique(rkey))
    n_edges = uk.numel() // 2
    assert v.shape[0] - n_edges + f.shape[0] == 2                                      # Euler characteristic of a sphere
    assert torch.equal(torch.unique(fl), torch.arange(v.shape[0], device="cuda"))      # every vertex is referenced
    p0, p1, p2 = (v[fl[:, k]].double() for k in range(3))
    cr = torch.cross(p1 - p0, p2 - p0, dim=-1)
    area = 0.5 * cr.norm(dim=1).sum().item() * h * h
    vol_m = (p0 * cr).sum().item() / 6.0 * h ** 3
    assert abs(area - 4 * np.pi * 0.36) < 1e-3 * 4 * np.pi * 0.36
    assert abs(vol_m - 4 / 3 * np.pi * 0.216) < 1e-3 * 4 / 3 * np.pi * 0.216          # positive: normals point outward
    rad = (v.double() * h - 1).norm(dim=1)
    assert (rad - 0.6).abs().max().item() < h * h
    p = str(tmp_path / "mesh_0.ply")
    M.write_ply(p, v * h - 1, f)
    v2, f2 = M.read_ply(p)
    assert np.array_equal(f2, f.cpu().numpy()) and np.array_equal(v2, (v * h - 1).cpu().numpy())


def test_export_stage0_mesh_from_a_trainer(tmp_path):
    """density volume -> marching cubes -> world coordinates -> mesh_0.ply (renderer.py:471-531,543-544) on the analytic occupancy"""
    from nerf2mesh_b200 import synthetic as S
    from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer
    tr = Stage0Trainer(Stage0Config(bound=1.0, num_rays=256, max_samples=256 * 128), seed=0)
    grid, bits, bricks = S.occupancy_regime("converged")
    tr.set_occupancy(bits, grid)
    thr = 0.5 * float(grid[grid > 0].min().item()) if (grid > 0).any() else 0.5
    v, f = M.export_stage0_mesh(tr, str(tmp_path), resolution=128, density_thresh=thr)
    assert f.shape[0] > 1000 and v.abs().max().item() <= 1.0 + 1e-6
    v2, f2 = M.read_ply(str(tmp_path / "mesh_0.ply"))
    assert v2.shape[0] == v.shape[0] and f2.shape[0] == f.shape[0]
    # the surface encloses exactly the occupied cells' region: its (outward-oriented) volume is positive and below the cube's
    p0, p1, p2 = (v[f.long()[:, k]].double() for k in range(3))
    vol_m = (p0 * torch.cross(p1 - p0, p2 - p0, dim=-1)).sum().item() / 6.0
    assert 0 < vol_m < 8.0
