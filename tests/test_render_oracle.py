"""CPU: the alive-ray round renderer (oracle/render_oracle.py = the algorithm of csrc/render.cu + Stage0Trainer.render) against the
all-samples forward of the train oracle: same image, fewer sample rows once rays terminate, capacity clipping and continuation rounds."""
import numpy as np
import torch

from nerf2mesh_b200 import synthetic as S
from oracle import render_oracle as RO
from oracle import train_oracle as T


def _scene(N=40, dense=True):
    torch.manual_seed(0)
    f = T.OracleField(1.0)
    if dense:          # an opaque medium (sigma = 300: alpha ~ 0.64 per sample), so that rays stop after about nine samples
        inner = f

        class Opaque(torch.nn.Module):
            def forward(self, x, d, shading="full", amp=True):
                s, c, sp = inner(x, d, shading, amp)
                return torch.full_like(s, 300.0), c, sp
        f = Opaque()
    grid, bits, bricks = S.occupancy_regime("converged")
    g = torch.Generator().manual_seed(1)
    ro, rd, _, _ = S.sample_rays(S.orbit_cameras(100), S.lego_intrinsics(), 800, 800, N, g)
    bg = torch.rand(N, 3, generator=g)
    return f, ro, rd, bits, bg


def test_rounds_equal_all_samples_and_stop_early():
    f, ro, rd, bits, bg = _scene()
    cfg = dict(bound=1.0, C=1, H=128, ref_composite=True)
    with torch.no_grad():
        full = T.render_train(f, ro, rd, bits, cfg, torch.zeros(ro.shape[0]), bg, "full", True)
    out = RO.render_rounds(f, ro, rd, bits, cfg, bg, capacity=ro.shape[0] * 512)            # no width is clipped
    assert (full["weights_sum"] > 0.99).float().mean().item() > 0.5            # the medium really is opaque for most rays
    assert (out["image"] - full["image"]).abs().max().item() <= 2e-4
    assert (out["weights_sum"] - full["weights_sum"]).abs().max().item() <= 2e-4
    assert (out["depth"] - full["depth"]).abs().max().item() <= 2e-4 * max(1.0, full["depth"].abs().max().item())
    assert out["rounds"] == len(RO.SCHEDULE)
    # early termination: the rounds evaluate far fewer rows than the march emits for the whole batch, padding included
    assert out["rows"] < 0.5 * full["num_points"], (out["rows"], full["num_points"])


def test_capacity_clipping_and_continuation_give_the_same_image():
    f, ro, rd, bits, bg = _scene(N=24, dense=False)        # a thin medium: rays run to the end of the volume
    cfg = dict(bound=1.0, C=1, H=128, ref_composite=True)
    with torch.no_grad():
        full = T.render_train(f, ro, rd, bits, cfg, torch.zeros(ro.shape[0]), bg, "full", True)
    a = RO.render_rounds(f, ro, rd, bits, cfg, bg, capacity=24 * 512)
    b = RO.render_rounds(f, ro, rd, bits, cfg, bg, schedule=(2, 2), capacity=24 * 3, more=64)      # clipped widths, many continuation rounds
    for out in (a, b):
        assert (out["image"] - full["image"]).abs().max().item() <= 2e-4
        assert (out["weights_sum"] - full["weights_sum"]).abs().max().item() <= 2e-4
    # without termination the doubling schedule pays for its few rounds with padded rows (a ray that needs 70 more samples gets a slab
    # of 128): 1.4x the marched samples on this batch, 1.25x with twelve rounds of (8, 8, 16, 16, 32, 32, ...), 1.15x with 64 rounds of 16
    assert full["num_points"] <= a["rows"] <= 1.6 * full["num_points"]
    assert b["rounds"] > 2 + 5                                       # the short schedule did not finish: further rounds ran
    assert a["rounds"] == len(RO.SCHEDULE)
