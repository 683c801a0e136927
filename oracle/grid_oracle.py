"""oracle/grid_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (PyTorch, vectorised over samples) of the reference's multiresolution grid
encoder: forward with optional dy_dx, table-gradient scatter, input gradient, TV gradient.
Follows gridencoder/src/gridencoder.cu (cited per function) and gridencoder/grid.py.
fp16 tables are emulated with the at::Half operator semantics the reference kernel has
(product rounded to half, sum rounded to half; torch/headeronly/util/Half.h:501-531).
"""
import numpy as np
import torch

PRIMES = [1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737]   # gridencoder.cu:54
U32 = 0xFFFFFFFF


def level_offsets(input_dim=3, num_levels=16, per_level_scale=2.0, base_resolution=16, log2_hashmap_size=19,
                  align_corners=False):
    """grid.py:124-134."""
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32)


def _level_geom(level, S, H, offsets):
    """gridencoder.cu:137-139 (exp2f in float32)."""
    rows = int(offsets[level + 1] - offsets[level])
    scale = np.float32(np.exp2(np.float32(level) * np.float32(S))) * np.float32(H) - np.float32(1.0)
    scale = np.float32(scale)
    res = int(np.ceil(scale)) + 1
    return scale, res, rows


def _row_index(pg, res, rows, gridtype, align_corners):
    """gridencoder.cu:66-84.  pg: int64 [B, D] lattice coords (uint32 semantics)."""
    D = pg.shape[1]
    stride, idx = 1, torch.zeros(pg.shape[0], dtype=torch.int64)
    d_used = 0
    for d in range(D):
        if stride <= rows:
            idx = (idx + pg[:, d] * stride) & U32
            stride = (stride * (res if align_corners else res + 1)) & U32
            d_used += 1
    if gridtype == 0 and stride > rows:
        h = torch.zeros(pg.shape[0], dtype=torch.int64)
        for d in range(D):
            h = h ^ ((pg[:, d] * PRIMES[d]) & U32)
        idx = h
    return idx % rows


def _smooth(v):
    return v * v * (3.0 - 2.0 * v)


def _smooth_d(v):
    return 6 * v * (1.0 - v)


def _h(x):
    """round to fp16 and back (at::Half storage)."""
    return x.to(torch.float16).to(torch.float32)


def grid_encode_forward(inputs, embeddings, offsets, S, H, max_level=None, gridtype=0, align_corners=False,
                        interp=0, calc_dy_dx=False):
    """gridencoder.cu:88-244.  inputs [B,D] float32 in [0,1]; embeddings [rows,C] float32 or float16.
    Returns outputs [L,B,C] (kernel layout) in the table dtype and dy_dx [B, L*D*C] or None."""
    inputs = inputs.float()
    B, D = inputs.shape
    L = len(offsets) - 1
    C = embeddings.shape[1]
    half = embeddings.dtype == torch.float16
    emb = embeddings.float()
    max_level = L if max_level is None else min(max_level, L)
    out = torch.zeros(L, B, C)
    dy_dx = torch.zeros(B, L, D, C) if calc_dy_dx else None
    oob = ((inputs < 0) | (inputs > 1)).any(-1)
    for level in range(max_level):
        scale, res, rows = _level_geom(level, S, H, offsets)
        tab = emb[int(offsets[level]):int(offsets[level + 1])]
        # pos = x*scale + 0.5 is an FMA in the compiled kernel: emulate through float64
        pos = (inputs.double() * float(scale) + (0.0 if align_corners else 0.5)).float()
        pg = torch.floor(pos).clamp(min=0).to(torch.int64)
        frac = pos - pg.float()
        dfrac = torch.ones_like(frac)
        if interp == 1:
            dfrac = _smooth_d(frac)
            frac = _smooth(frac)
        acc = torch.zeros(B, C)
        for corner in range(1 << D):
            w = torch.ones(B)
            p = pg.clone()
            for d in range(D):
                if corner & (1 << d):
                    w = w * frac[:, d]; p[:, d] += 1
                else:
                    w = w * (1 - frac[:, d])
            row = _row_index(p, res, rows, gridtype, align_corners)
            g = tab[row]
            if half:
                acc = _h(acc + _h(w[:, None] * g))
            else:
                acc = (acc.double() + w[:, None].double() * g.double()).float()     # FFMA
        acc[oob] = 0
        out[level] = acc
        if calc_dy_dx:
            for gd in range(D):
                gacc = torch.zeros(B, C)
                others = [d for d in range(D) if d != gd]
                for corner in range(1 << (D - 1)):
                    w = torch.full((B,), float(scale))
                    p = pg.clone()
                    for nd, d in enumerate(others):
                        if corner & (1 << nd):
                            w = w * frac[:, d]; p[:, d] += 1
                        else:
                            w = w * (1 - frac[:, d])
                    lo = tab[_row_index(p, res, rows, gridtype, align_corners)]
                    p[:, gd] += 1
                    hi = tab[_row_index(p, res, rows, gridtype, align_corners)]
                    if half:
                        diff = _h(hi - lo)
                        gacc = _h(gacc + _h(w[:, None] * diff * dfrac[:, gd:gd + 1]))
                    else:
                        gacc = gacc + w[:, None] * (hi - lo) * dfrac[:, gd:gd + 1]
                gacc[oob] = 0
                dy_dx[:, level, gd] = gacc
    dt = embeddings.dtype
    return out.to(dt), (dy_dx.reshape(B, L * D * C).to(dt) if calc_dy_dx else None)


def grid_encode_backward(grad, inputs, offsets, n_rows, S, H, max_level=None, gridtype=0, align_corners=False,
                         interp=0, dy_dx=None):
    """gridencoder.cu:248-368.  grad [L,B,C] -> grad_embeddings [rows,C] (float64 accumulation = the
    order-independent sum the atomics approximate) and grad_inputs [B,D] or None."""
    inputs = inputs.float()
    B, D = inputs.shape
    L = len(offsets) - 1
    C = grad.shape[2]
    max_level = L if max_level is None else min(max_level, L)
    gemb = torch.zeros(n_rows, C, dtype=torch.float64)
    oob = ((inputs < 0) | (inputs > 1)).any(-1)
    g = grad.double().clone()
    g[:, oob] = 0
    for level in range(max_level):
        scale, res, rows = _level_geom(level, S, H, offsets)
        pos = (inputs.double() * float(scale) + (0.0 if align_corners else 0.5)).float()
        pg = torch.floor(pos).clamp(min=0).to(torch.int64)
        frac = pos - pg.float()
        if interp == 1:
            frac = _smooth(frac)
        for corner in range(1 << D):
            w = torch.ones(B)
            p = pg.clone()
            for d in range(D):
                if corner & (1 << d):
                    w = w * frac[:, d]; p[:, d] += 1
                else:
                    w = w * (1 - frac[:, d])
            row = _row_index(p, res, rows, gridtype, align_corners) + int(offsets[level])
            gemb.index_add_(0, row, w[:, None].double() * g[level])
    ginp = None
    if dy_dx is not None:
        dd = dy_dx.double().reshape(B, L, D, C)
        ginp = torch.einsum("lbc,bldc->bd", grad.double(), dd)
    return gemb, ginp


def grad_total_variation(inputs, embeddings, offsets, weight, S, H, gridtype=0, align_corners=False):
    """gridencoder.cu:506-609 (fp32).  Returns the TV gradient to ADD to embeddings.grad (float64)."""
    inputs = inputs.float()
    B, D = inputs.shape
    L = len(offsets) - 1
    C = embeddings.shape[1]
    emb = embeddings.float()
    out = torch.zeros(emb.shape[0], C, dtype=torch.float64)
    oob = ((inputs < 0) | (inputs > 1)).any(-1)
    w = np.float32(weight) / np.float32(2 * D)
    for level in range(L):
        scale, res, rows = _level_geom(level, S, H, offsets)
        tab = emb[int(offsets[level]):int(offsets[level + 1])]
        pos = (inputs.double() * float(scale) + (0.0 if align_corners else 0.5)).float()
        pg = torch.floor(pos).clamp(min=0).to(torch.int64)
        row = _row_index(pg, res, rows, gridtype, align_corners)
        centre = tab[row]
        s = torch.zeros(B, C); sq = torch.zeros(B, C)
        for d in range(D):
            cur = pg[:, d]
            for delta, ok in ((1, cur < res), (-1, cur > 0)):
                p = pg.clone(); p[:, d] = (cur + delta).clamp(min=0)
                nb = tab[_row_index(p, res, rows, gridtype, align_corners)]
                dv = (centre - nb) * ok[:, None].float()
                s = s + dv; sq = sq + dv * dv
        val = float(w) * s * torch.rsqrt(sq + 1e-9)
        val[oob] = 0
        out.index_add_(0, row + int(offsets[level]), val.double())
    return out
