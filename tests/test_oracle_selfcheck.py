"""CPU: internal consistency of the oracle pieces (no golden data needed)."""
import numpy as np
import torch

import cases
from oracle import raymarching_oracle as R, sh_oracle


def test_sh_first_bands_match_reference_table():
    """The first 9 polynomials as the reference spells them out (shencoder.cu:50-60)."""
    v = cases.sh_case().numpy().astype(np.float64)
    x, y, z = v.T
    o, _ = sh_oracle.sh_encode(v, 3)
    exp = np.stack([np.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z,
                    -0.48860251190291987 * x, 1.0925484305920792 * x * y, -1.0925484305920792 * y * z,
                    0.94617469575755997 * z * z - 0.31539156525251999, -1.0925484305920792 * x * z,
                    0.54627421529603959 * (x * x - y * y)], -1)
    assert np.abs(o - exp).max() < 1e-12


def test_composite_backward_is_the_gradient_of_forward():
    """True for grad_weights == 0.  (The reference folds the per-sample grad_weights into the weights_sum
    term, raymarching.cu:676, which is NOT the exact gradient of sum_j g_j w_j -- the kernels reproduce the
    reference formula, the autograd oracle is only comparable without that term.)"""
    from oracle.train_oracle import composite_train
    c = cases.composite_case(N=24, max_cnt=40, sigma_scale=15.0)
    c["grad_weights"] = torch.zeros_like(c["grad_weights"])
    sig = c["sigmas"].double().requires_grad_(True); rgb = c["rgbs"].double().requires_grad_(True)
    w, ws, d, im = composite_train(sig.float(), rgb.float(), c["ts"], c["rays"], 1e-4)
    (ws * c["grad_weights_sum"]).sum().add((im * c["grad_image"]).sum()).add((d * c["grad_depth"]).sum()).add((w * c["grad_weights"]).sum()).backward()
    w0, ws0, d0, im0 = R.composite_rays_train_forward(c["sigmas"].numpy(), c["rgbs"].numpy(), c["ts"].numpy(), c["rays"].numpy(), 1e-4)
    gs, gr = R.composite_rays_train_backward(c["grad_weights"].numpy(), c["grad_weights_sum"].numpy(), c["grad_depth"].numpy(),
                                             c["grad_image"].numpy(), c["sigmas"].numpy(), c["rgbs"].numpy(), c["ts"].numpy(),
                                             c["rays"].numpy(), ws0, d0, im0, 1e-4)
    assert np.abs(gs - sig.grad.numpy()).max() <= 1e-3 * np.abs(gs).max()
    assert np.abs(gr - rgb.grad.numpy()).max() <= 1e-3 * np.abs(gr).max()


def test_train_oracle_step_decreases_loss():
    from nerf2mesh_b200 import synthetic as S
    from oracle import train_oracle as T
    torch.manual_seed(0)
    f = T.OracleField(1.0)
    opt = torch.optim.Adam(f.parameters(), lr=1e-2, eps=1e-15)
    ro, rd = cases.rays(48, 0)
    grid, bits, bricks = S.occupancy_regime("converged")
    gt = S.render_bricks(ro, rd, bricks)
    cfg = dict(bound=1.0, C=1, H=128)
    bg = torch.rand(48, 3)
    losses = [T.train_step(f, opt, ro, rd, gt, bits, cfg, torch.zeros(48), bg)[0] for _ in range(3)]
    assert losses[-1] < losses[0]


def test_reference_backward_compositor_matches_autograd_without_weight_losses():
    """oracle/train_oracle._CompositeRef (reference backward formula, used when the entropy regulariser is on) gives the
    autograd gradients of the padded-tensor compositor as long as no loss touches `weights`; with lambda_entropy > 0 it
    follows the reference's folding of grad_weights (raymarching.cu:676) and therefore differs from autograd."""
    from nerf2mesh_b200 import synthetic as S
    from oracle import train_oracle as T
    torch.manual_seed(0)
    ro, rd = cases.rays(24, 1)
    grid, bits, bricks = S.occupancy_regime("converged")
    gt = S.render_bricks(ro, rd, bricks)
    bg = torch.rand(24, 3)
    grads = {}
    for ref in (False, True):
        for lam in (0.0, 1e-2):
            torch.manual_seed(1)
            f = T.OracleField(1.0)
            with torch.no_grad():
                f.sigma_net.net[1].weight.mul_(30.0)           # some opacity, so that rays terminate and weights spread
            cfg = dict(bound=1.0, C=1, H=128, ref_composite=ref)
            out = T.render_train(f, ro, rd, bits, cfg, torch.zeros(24), bg, "full", True)
            loss = T.train_loss(out, gt, bg, 0.1, 1e-5, lambda_entropy=lam)
            loss.backward()
            grads[(ref, lam)] = (loss.item(), f.sigma_net.net[0].weight.grad.clone(), f.color_net.net[0].weight.grad.clone())
    for i in (1, 2):
        a, b = grads[(False, 0.0)][i], grads[(True, 0.0)][i]
        assert (a - b).abs().max() <= 2e-3 * a.abs().max()
    assert abs(grads[(False, 0.0)][0] - grads[(True, 0.0)][0]) <= 1e-5 * abs(grads[(False, 0.0)][0])
    assert abs(grads[(False, 1e-2)][0] - grads[(True, 1e-2)][0]) <= 1e-5 * abs(grads[(True, 1e-2)][0])      # same loss value
    assert grads[(True, 1e-2)][0] > grads[(True, 0.0)][0]
    d = (grads[(True, 1e-2)][1] - grads[(True, 0.0)][1]).abs().max()
    assert d > 0                                                                                      # the term reaches sigma_net
