"""CPU: host-side logic and the C ABI surface (no kernel is launched)."""
import ctypes
import os
import re

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("n2m_b200.h", "n2m_b200_fused.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(n2m_[a-zA-Z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from nerf2mesh_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_lib.lib, s), f"libn2m_b200.so does not export {s}"
    assert _lib.lib.n2m_version() == 100
    assert _lib.last_error() == ""


def test_bindings_cover_the_headers():
    from nerf2mesh_b200 import _lib, parallel, sampler, stage0  # noqa: F401  (register the fused / data-parallel signatures)
    bound = set(_lib.SIGNATURES) | {"n2m_last_error", "n2m_version", "n2m_launch_count", "n2m_s0_wpack_bytes",
                                   "n2m_s0_mlp_param_count", "n2m_s0_init", "n2m_dp_ctx_bytes"}
    assert set(declared_symbols()) <= bound, set(declared_symbols()) - bound


def test_params_struct_matches_header():
    from nerf2mesh_b200.stage0 import S0Params
    hdr = open(os.path.join(ROOT, "include", "n2m_b200_fused.h")).read()
    body = hdr[hdr.index("typedef struct {"):hdr.index("} n2m_s0_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(float|uint32_t)\s+([a-z0-9_A-Z]+);", body)
    assert [n for _, n in fields] == [n for n, _ in S0Params._fields_]
    assert ctypes.sizeof(S0Params) == 4 * len(fields)


def test_morton_spread_equals_reference_form():
    def ref(v):
        v = (v * 0x00010001) & 0xFF0000FF; v = (v * 0x00000101) & 0x0F00F00F
        v = (v * 0x00000011) & 0xC30C30C3; v = (v * 0x00000005) & 0x49249249
        return v & 0xFFFFFFFF

    def ours(v):
        v &= 0x7FF
        v = (v | (v << 16)) & 0x070000FF; v = (v | (v << 8)) & 0x0700F00F
        v = (v | (v << 4)) & 0x430C30C3; v = (v | (v << 2)) & 0x49249249
        return v & 0xFFFFFFFF
    assert all(ref(v) == ours(v) for v in range(2048))
    from oracle import raymarching_oracle as R
    c = np.random.default_rng(0).integers(0, 128, (1000, 3))
    m = R.morton3D(c)
    assert np.array_equal(R.morton3D_invert(m), c)
    from nerf2mesh_b200.synthetic import _morton_np
    assert np.array_equal(_morton_np(c), m.astype(np.int64))


def test_level_offsets_match_oracle_and_survey():
    from nerf2mesh_b200.gridencoder.grid import GridEncoder, level_offsets
    from oracle import grid_oracle
    for bound, total in ((1, 6119864), (16, 6837544)):
        pls = float(np.exp2(np.log2(2048 * bound / 16) / 15))
        a = level_offsets(3, 16, pls, 16, 19, False)
        b = grid_oracle.level_offsets(3, 16, pls, 16, 19, False)
        assert np.array_equal(a, b) and int(a[-1]) == total
    enc = GridEncoder(level_dim=2, desired_resolution=2048)
    assert enc.embeddings.shape == (6119864, 2) and enc.output_dim == 32
    assert enc.embeddings.abs().max() <= 1e-4
    assert "GridEncoder" in repr(enc)


def test_packbits_oracle_and_synthetic_agree():
    from nerf2mesh_b200 import synthetic as S
    from oracle import raymarching_oracle as R
    g = torch.rand(2, 4096)
    assert np.array_equal(S.packbits_host(g, 0.3).numpy(), R.packbits(g.numpy(), 0.3))
    grid, bits, _ = S.occupancy_regime("converged", H=32)
    assert bits.numel() == 32 ** 3 // 8 and 0.02 < grid.mean() < 0.5


def test_stage0_config_mirrors_renderer():
    from nerf2mesh_b200.stage0 import Stage0Config
    c = Stage0Config(bound=16.0)
    assert c.cascade == 5 and abs(c.per_level_scale - 1.662476) < 1e-5
    c = Stage0Config(bound=4.0, contract=True)
    assert c.bound == 2.0 and c.cascade == 2 and c.real_bound == 4.0              # renderer.py:74-82
    assert Stage0Config(num_rays=100, max_samples=1000).max_samples == 1024


def test_drop_in_module_names():
    import nerf2mesh_b200
    rm, ge, sh = nerf2mesh_b200.install()
    import gridencoder
    import raymarching
    import shencoder
    for n in ("near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "flatten_rays",
              "march_rays_train", "composite_rays_train", "march_rays", "composite_rays"):
        assert callable(getattr(raymarching, n))
    assert gridencoder.GridEncoder is ge.GridEncoder and shencoder.SHEncoder is sh.SHEncoder


def test_peer_adam_slices_cover_rows_and_stay_aligned():
    """PeerAdam row slices: disjoint cover of the table, slice length a multiple of 4 so that the float2 colour
    moments stored behind the density moments are 8-byte aligned (6119864 rows / 8 ranks = 764983 is odd)."""
    from nerf2mesh_b200.parallel import slice_rows
    from nerf2mesh_b200.gridencoder.grid import level_offsets
    rows = int(level_offsets(3, 16, float(np.exp2(np.log2(2048 / 16) / 15)), 16, 19, False)[-1])
    for R in (rows, 1, 7, 1000003):
        for W in range(1, 9):
            per = slice_rows(R, W)
            assert per % 4 == 0 and per * W >= R
            lo = [min(R, r * per) for r in range(W)]
            hi = [min(R, (r + 1) * per) for r in range(W)]
            assert lo[0] == 0 and hi[-1] == R and all(hi[r] == lo[r + 1] for r in range(W - 1))
