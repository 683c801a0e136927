"""Stage-1 mesh operators with the call surface of `nvdiffrast.torch` as the reference uses it (nerf/renderer.py:126-128, 856-901):

    glctx = RasterizeCudaContext()
    rast, rast_db = rasterize(glctx, pos, tri, (h, w))          # pos [1,V,4] clip space, tri [F,3] int32 -> rast [1,h,w,4]
    out, out_db   = interpolate(attr, rast, tri)                 # attr [1,V,A] -> [1,h,w,A], differentiable w.r.t. attr

over the sm_100a kernels of csrc/raster.cu (C ABI: include/n2m_b200_raster.h).  `rast[..., :] = (u, v, z/w, triangle_id + 1)`.
Not provided: image-space derivative outputs (`rast_db`, `out_db` are None: the reference ignores them, renderer.py:860-863),
gradients of rasterize w.r.t. vertex positions, `antialias` (renderer.py:886-887) -- see DESIGN.md "stage 1".
No CPU fallback: tensors must live on a CUDA device.
"""
import torch

from . import _lib
from ._lib import P, U, call, ptr, stream

_lib.register({
    "n2m_rasterize": [P, U, P, U, U, U, P, P, P, P],
    "n2m_interpolate_forward": [P, U, U, P, P, U, P, P],
    "n2m_interpolate_backward": [P, P, P, U, U, U, P, P],
    "n2m_compact_covered": [P, P, P, U, U, P, P, P, P, P],
})


class RasterizeCudaContext:
    """Scratch owner (visibility buffer + large-triangle queue), the counterpart of dr.RasterizeCudaContext / RasterizeGLContext."""

    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        self._vis = None
        self._queue = None

    def scratch(self, num_pixels, num_tris):
        if self._vis is None or self._vis.numel() < num_pixels:
            self._vis = torch.empty(num_pixels, dtype=torch.int64, device=self.device)
        if self._queue is None or self._queue.numel() < num_tris + 1:
            self._queue = torch.empty(num_tris + 1, dtype=torch.int32, device=self.device)
        return self._vis, self._queue


RasterizeGLContext = RasterizeCudaContext        # the reference picks either (renderer.py:126-128); both map to the CUDA kernels


def rasterize(glctx, pos, tri, resolution, ranges=None, grad_db=True):
    """dr.rasterize: pos [1,V,4] float32 clip space (or [V,4]), tri [F,3] int32, resolution (h, w) -> (rast [1,h,w,4], None)."""
    if ranges is not None:
        raise NotImplementedError("range mode is not used by the reference")
    if not pos.is_cuda:
        raise RuntimeError("rasterize: pos must be a CUDA tensor (nerf2mesh_b200 has no CPU path)")
    if pos.dim() == 3:
        if pos.shape[0] != 1:
            raise NotImplementedError("instanced mode with minibatch > 1 is not used by the reference")
        pos = pos[0]
    pos = pos.detach().float().contiguous()
    tri = tri.int().contiguous()
    h, w = int(resolution[0]), int(resolution[1])
    vis, queue = glctx.scratch(h * w, tri.shape[0])
    rast = torch.empty(1, h, w, 4, device=pos.device, dtype=torch.float32)
    call("n2m_rasterize", ptr(pos), pos.shape[0], ptr(tri), tri.shape[0], h, w, ptr(vis), ptr(queue), ptr(rast), stream())
    return rast, None


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attr, rast, tri):
        V, A = attr.shape
        n = rast.shape[0] * rast.shape[1] * rast.shape[2]
        out = torch.empty(*rast.shape[:3], A, device=attr.device, dtype=torch.float32)
        call("n2m_interpolate_forward", ptr(attr), V, A, ptr(rast), ptr(tri), n, ptr(out), stream())
        ctx.save_for_backward(rast, tri)
        ctx.dims = (V, A, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        rast, tri = ctx.saved_tensors
        V, A, n = ctx.dims
        g = torch.zeros(V, A, device=grad_out.device, dtype=torch.float32)
        call("n2m_interpolate_backward", ptr(grad_out.float().contiguous()), ptr(rast), ptr(tri), n, V, A, ptr(g), stream())
        return g, None, None


def interpolate(attr, rast, tri, rast_db=None, diff_attrs=None):
    """dr.interpolate: attr [1,V,A] (or [V,A]) float32, rast [1,h,w,4], tri [F,3] -> (out [1,h,w,A], None)."""
    if not attr.is_cuda:
        raise RuntimeError("interpolate: attr must be a CUDA tensor")
    a = attr[0] if attr.dim() == 3 else attr
    if a.shape[-1] < 1 or a.shape[-1] > 4:
        raise RuntimeError("interpolate: 1..4 attributes per vertex are supported")
    out = _Interpolate.apply(a.float().contiguous(), rast.contiguous(), tri.int().contiguous())
    return out, None


def compact_covered(rast, xyz, dirs, cap=None):
    """Covered pixels of `rast` ([1,h,w,4]) with their interpolated positions `xyz` [h*w,3] and view directions `dirs` [h*w,3]:
    -> (count [1] int32 on the device, pixel index [cap] int32, points [cap,3], dirs [cap,3]); no host synchronisation."""
    n = rast.shape[1] * rast.shape[2]
    cap = int(cap or n)
    dev = rast.device
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    pix = torch.empty(cap, dtype=torch.int32, device=dev)
    pts = torch.zeros(cap, 3, device=dev); pd = torch.zeros(cap, 3, device=dev)
    call("n2m_compact_covered", ptr(rast), ptr(xyz.contiguous()), ptr(dirs.contiguous()), n, cap, ptr(counter), ptr(pix), ptr(pts), ptr(pd), stream())
    return counter, pix, pts, pd
