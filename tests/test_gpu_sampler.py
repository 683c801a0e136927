"""GPU: on-device batch sampling (n2m_s0_gen_rays) against the torch restatement of get_rays / collate
(nerf/utils.py:236-290, nerf/provider.py:300-331) that nerf2mesh_b200.synthetic.sample_rays uses on the host."""
import pytest
import torch

from nerf2mesh_b200 import synthetic as S
from nerf2mesh_b200.sampler import DeviceRaySampler

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [4, 3, 0])
def test_device_sampler_matches_get_rays(C):
    B, H, W, N = 7, 60, 80, 5000
    poses = S.orbit_cameras(B, seed=3)
    intr = torch.from_numpy(S.lego_intrinsics()) * torch.tensor([0.1, 0.1, 0.1, 0.075])
    g = torch.Generator().manual_seed(0)
    images = torch.randint(0, 256, (B, H, W, C), dtype=torch.uint8, generator=g) if C else None
    img = torch.randint(0, B, (N,), generator=g); pix = torch.randint(0, H * W, (N,), generator=g)
    pix[:4] = torch.tensor([0, W - 1, H * W - W, H * W - 1])                 # image corners
    smp = DeviceRaySampler(poses, intr, H, W, images)
    ro, rd, gt = smp.sample(N, img.cuda(), pix.cuda())
    torch.cuda.synchronize()
    fx, fy, cx, cy = [float(v) for v in intr]
    i = (pix % W).float() + 0.5; j = (pix // W).float() + 0.5
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    rd_ref = (dirs.unsqueeze(1) @ poses[img, :3, :3].transpose(-1, -2)).squeeze(1)          # utils.py:282
    assert torch.equal(ro.cpu(), poses[img, :3, 3])
    assert (rd.cpu() - rd_ref).abs().max().item() <= 2e-6 * rd_ref.abs().max().item()
    if C:
        assert gt.shape == (N, C)
        assert torch.equal(gt.cpu(), images[img, pix // W, pix % W].float() / 255)          # provider.py:321
    else:
        assert gt is None
    # indices drawn on the device stay in range; writing into caller buffers works (a trainer slot)
    a, b = smp.draw_indices(N)
    assert a.min() >= 0 and a.max() < B and b.min() >= 0 and b.max() < H * W
    out = (torch.zeros(N, 3, device="cuda"), torch.zeros(N, 3, device="cuda"), torch.zeros(N, C, device="cuda") if C else None)
    smp.sample(N, img.cuda(), pix.cuda(), out=out)
    assert torch.equal(out[0], ro) and torch.equal(out[1], rd)
