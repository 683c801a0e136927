"""CPU: the generated marching-cubes case table (nerf2mesh_b200/mc_table.py) and the oracle built on it (oracle/mcubes_oracle.py):
properties any correct marching cubes has -- PyMCubes itself cannot be run here (parity unpinned, see the oracle's header)."""
import numpy as np

from nerf2mesh_b200 import mc_table as T
from nerf2mesh_b200 import mesh as M
from oracle import mcubes_oracle as O


def _grid(R):
    g = np.linspace(-1, 1, R)
    return np.meshgrid(g, g, g, indexing="ij")


def test_table_is_complete_and_consistent():
    import hashlib
    # the generator is deterministic: the kernels, the oracle and the committed GPU results all refer to THIS table
    assert hashlib.sha1(T.TRI_TABLE.tobytes()).hexdigest() == "38c7da8266ca550630c5abb3a0ca1ac7415d2a98"
    assert T.TRI_TABLE.shape == (256, 16) and T.NUM_TRIS[0] == 0 and T.NUM_TRIS[255] == 0
    assert T.NUM_TRIS.max() == 5 and T.NUM_TRIS.sum() == 820            # as many triangles as the classic Lorensen-Cline table
    for mask in range(256):
        row = T.TRI_TABLE[mask]
        used = row[row >= 0]
        assert len(used) == 3 * T.NUM_TRIS[mask] and (row[len(used):] == -1).all()
        # exactly the crossed edges appear
        crossed = {e for e, (a, b) in enumerate(T.EDGE_CORNERS) if ((mask >> a) & 1) != ((mask >> b) & 1)}
        assert set(int(e) for e in used) == crossed
        # within the cell every polygon edge that is not on a cube face is shared by two triangles; face segments appear once
        # (checked globally by the closedness of whole meshes below)
    # a single inside corner gives one triangle whose normal points away from that corner
    for c in range(8):
        tri = T.TRI_TABLE[1 << c][:3]
        mids = np.array([(np.array(T.corner_offset(int(T.EDGE_CORNERS[e, 0]))) + np.array(T.corner_offset(int(T.EDGE_CORNERS[e, 1])))) / 2 for e in tri])
        n = np.cross(mids[1] - mids[0], mids[2] - mids[0])
        assert np.dot(n, mids.mean(0) - np.array(T.corner_offset(c), float)) > 0


def test_every_configuration_closes_against_its_neighbours():
    """random binary volumes exercise all 256 cases including both resolutions of every ambiguous face: the surface must stay a closed,
    consistently oriented 2-manifold-with-shared-edges (each edge used exactly twice, once per direction)"""
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(6):
        vol = np.zeros((9, 9, 9))
        vol[1:-1, 1:-1, 1:-1] = rng.random((7, 7, 7)) + (rng.random((7, 7, 7)) > 0.5)          # values away from the iso level 1.0
        v, f = O.marching_cubes(vol, 1.0)
        pr = O.mesh_properties(v, f)
        assert pr["closed"] and pr["consistent"], (trial, pr)
        assert pr["volume"] > 0
        inside = vol > 1.0
        case = np.zeros((8, 8, 8), np.int64)
        for c in range(8):
            ox, oy, oz = T.corner_offset(c)
            case |= inside[ox:8 + ox, oy:8 + oy, oz:8 + oz].astype(np.int64) << c
        seen |= set(case.reshape(-1).tolist())
    assert len(seen) > 200, len(seen)


def test_sphere_torus_and_two_components():
    R = 40
    x, y, z = _grid(R)
    h = 2.0 / (R - 1)
    # sphere of radius 0.6: inside = positive
    v, f = O.marching_cubes(0.6 - np.sqrt(x * x + y * y + z * z), 0.0)
    pr = O.mesh_properties(v, f)
    assert pr["closed"] and pr["consistent"] and pr["euler"] == 2 and pr["degenerate"] == 0
    assert abs(pr["area"] * h * h - 4 * np.pi * 0.36) < 0.01 * 4 * np.pi * 0.36
    assert abs(pr["volume"] * h ** 3 - 4 / 3 * np.pi * 0.216) < 0.01 * 4 / 3 * np.pi * 0.216          # positive: normals point outward
    # vertices lie on the sphere to the interpolation error of a linear cut through a curved field
    rad = np.linalg.norm(v * h - 1, axis=1)
    assert np.abs(rad - 0.6).max() < 0.5 * h * h / 0.6 + 1e-9
    # every vertex sits on a grid edge: two integer coordinates
    frac = np.abs(v - np.round(v)) > 1e-12
    assert (frac.sum(1) <= 1).all()
    # torus: Euler characteristic 0
    q = np.sqrt(x * x + y * y) - 0.55
    v, f = O.marching_cubes(0.22 - np.sqrt(q * q + z * z), 0.0)
    pr = O.mesh_properties(v, f)
    assert pr["closed"] and pr["consistent"] and pr["euler"] == 0
    assert abs(pr["volume"] * h ** 3 - 2 * np.pi ** 2 * 0.55 * 0.22 ** 2) < 0.03 * 2 * np.pi ** 2 * 0.55 * 0.22 ** 2
    # two spheres: two components
    d1 = np.sqrt((x - 0.45) ** 2 + y * y + z * z); d2 = np.sqrt((x + 0.45) ** 2 + y * y + z * z)
    v, f = O.marching_cubes(np.maximum(0.3 - d1, 0.3 - d2), 0.0)
    pr = O.mesh_properties(v, f)
    assert pr["closed"] and pr["euler"] == 4
    # a density-like field with an iso level > 0 and a surface touching the volume boundary stays open only there
    v, f = O.marching_cubes(np.exp(-4 * (x * x + y * y)) * 20, 10.0)         # a cylinder along z, cut by the volume faces
    assert len(f) > 0 and not O.mesh_properties(v, f)["closed"]


def test_ply_round_trip(tmp_path):
    R = 12
    x, y, z = _grid(R)
    v, f = O.marching_cubes(0.6 - np.sqrt(x * x + y * y + z * z), 0.0)
    p = str(tmp_path / "mesh_0.ply")
    M.write_ply(p, v.astype(np.float32), f.astype(np.int32))
    v2, f2 = M.read_ply(p)
    assert np.array_equal(v2, v.astype(np.float32)) and np.array_equal(f2, f.astype(np.int32))
    head = open(p, "rb").read(200)
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\n") and b"property list uchar int vertex_indices" in head
