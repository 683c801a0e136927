"""Stage-0 -> stage-1 mesh hand-off on the device: marching cubes over the density volume + PLY writer.

Replaces the `mcubes.marching_cubes(sigmas, density_thresh)` call of NeRFRenderer.export_stage0 (nerf/renderer.py:526-529; PyMCubes is a
third-party CPU library) with the kernels of csrc/mcubes.cu (C ABI include/n2m_b200_mesh.h):

    vertices, triangles = marching_cubes(volume, isovalue)      # volume [X,Y,Z] float32 CUDA tensor
    # vertices [V,3] float32 in index coordinates (0 .. X-1), triangles [F,3] int32 -- the form PyMCubes returns (as tensors)

`export_stage0_mesh(trainer, path, resolution)` is the reference's export up to (not including) its CPU clean-up / decimation:
density volume (Stage0Trainer.density_volume == renderer.py:480-524) -> marching cubes -> `vertices / (resolution - 1) * 2 - 1` (:531) ->
`mesh_0.ply`.  No CPU fallback: the volume must live on a CUDA device.
"""
import os
import struct

import numpy as np
import torch

from . import _lib
from . import mc_table
from ._lib import F, P, U, call, ptr, stream

_lib.register({
    "n2m_mc_count": [P, U, U, U, F, P, P, P, P],
    "n2m_mc_emit": [P, U, U, U, F, P, P, P, P, P, P, P],
})

_tables = {}


def _device_tables(device):
    key = torch.device(device).index
    if key not in _tables:
        _tables[key] = (torch.from_numpy(mc_table.TRI_TABLE.copy()).to(device), torch.from_numpy(mc_table.NUM_TRIS.copy()).to(device))
    return _tables[key]


def marching_cubes(volume, isovalue):
    """volume [X,Y,Z] float32 CUDA tensor, isovalue float -> (vertices [V,3] float32 in index coordinates, triangles [F,3] int32).
    Inside = value > isovalue; normals point from inside to outside; vertices are shared between cells; deterministic order."""
    if not volume.is_cuda:
        raise RuntimeError("marching_cubes: the volume must be a CUDA tensor (nerf2mesh_b200 has no CPU path)")
    if volume.dim() != 3 or min(volume.shape) < 2:
        raise RuntimeError("marching_cubes: volume must be [X,Y,Z] with at least two samples per axis")
    vol = volume.float().contiguous()
    X, Y, Z = (int(s) for s in vol.shape)
    n = X * Y * Z
    dev = vol.device
    tri_table, num_tris = _device_tables(dev)
    vcount = torch.empty(n, dtype=torch.uint8, device=dev); tcount = torch.empty(n, dtype=torch.uint8, device=dev)
    call("n2m_mc_count", ptr(vol), X, Y, Z, float(isovalue), ptr(num_tris), ptr(vcount), ptr(tcount), stream())
    vinc = torch.cumsum(vcount, 0, dtype=torch.int32); tinc = torch.cumsum(tcount, 0, dtype=torch.int32)
    V, T = int(vinc[-1].item()), int(tinc[-1].item())                     # the one host read-back (output sizes)
    voff = vinc - vcount.to(torch.int32); toff = tinc - tcount.to(torch.int32)
    del vinc, tinc
    vertices = torch.empty(V, 3, device=dev); triangles = torch.empty(T, 3, dtype=torch.int32, device=dev)
    if V > 0 or T > 0:
        call("n2m_mc_emit", ptr(vol), X, Y, Z, float(isovalue), ptr(tri_table), ptr(tcount), ptr(voff), ptr(toff),
             ptr(vertices) if V > 0 else ptr(torch.empty(3, device=dev)), ptr(triangles) if T > 0 else ptr(torch.empty(3, dtype=torch.int32, device=dev)), stream())
    return vertices, triangles


def write_ply(path, vertices, triangles):
    """binary little-endian PLY (float32 x y z, uchar-counted int32 face lists): the format trimesh writes for `mesh_0.ply`
    (renderer.py:543-544) and reads back in stage 1"""
    v = np.ascontiguousarray(vertices.detach().cpu().numpy() if torch.is_tensor(vertices) else vertices, dtype="<f4")
    f = np.ascontiguousarray(triangles.detach().cpu().numpy() if torch.is_tensor(triangles) else triangles, dtype="<i4")
    header = ("ply\nformat binary_little_endian 1.0\n"
              f"element vertex {v.shape[0]}\nproperty float x\nproperty float y\nproperty float z\n"
              f"element face {f.shape[0]}\nproperty list uchar int vertex_indices\nend_header\n").encode("ascii")
    faces = np.empty(f.shape[0], dtype=[("n", "u1"), ("i", "<i4", (3,))])
    faces["n"] = 3; faces["i"] = f
    with open(path, "wb") as fh:
        fh.write(header); fh.write(v.tobytes()); fh.write(faces.tobytes())


def read_ply(path):
    """the inverse of write_ply (tests, and stage 1 picking the mesh up again)"""
    with open(path, "rb") as fh:
        data = fh.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    head = data[:end].decode("ascii").split("\n")
    nv = int(next(l for l in head if l.startswith("element vertex")).split()[-1])
    nf = int(next(l for l in head if l.startswith("element face")).split()[-1])
    v = np.frombuffer(data, dtype="<f4", count=3 * nv, offset=end).reshape(nv, 3)
    faces = np.frombuffer(data, dtype=[("n", "u1"), ("i", "<i4", (3,))], count=nf, offset=end + 12 * nv)
    assert (faces["n"] == 3).all()
    return v.copy(), faces["i"].copy()


def export_stage0_mesh(trainer, save_path, resolution=512, density_thresh=10.0):
    """NeRFRenderer.export_stage0 for the inner region up to its CPU post-processing (renderer.py:471-531,543-544): density volume ->
    marching cubes at min(mean_density, density_thresh) -> world coordinates -> `<save_path>/mesh_0.ply`.  Returns (vertices, triangles)
    on the device.  Cleaning / decimation (clean_mesh, decimate_mesh: pymeshlab) and the outer-region meshes stay the reference's code."""
    vol = trainer.density_volume(resolution=resolution, density_thresh=density_thresh)
    mean = getattr(trainer, "mean_density", None)
    thresh = min(float(mean.item()), density_thresh) if mean is not None else density_thresh
    v, f = marching_cubes(vol, thresh)
    v = v / (resolution - 1.0) * 2 - 1                      # renderer.py:531
    os.makedirs(save_path, exist_ok=True)
    write_ply(os.path.join(save_path, "mesh_0.ply"), v, f)
    return v, f
