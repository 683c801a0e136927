"""Cost of one lane of a spread global RED by payload (csrc/red_probe.cu): random 16-byte rows, table L2-resident (32 MB) or
larger than L2 (512 MB), 148 x 16 blocks x 256 threads x 64 REDs = 38.8 M lane-REDs (one train step's worth, unmerged).
    python profiles/redbench.py"""
import sys; sys.path.insert(0, '/root/repo')
import torch
from nerf2mesh_b200 import _lib
from nerf2mesh_b200._lib import I, P, U, call, ptr, stream
from profiles.probes import call as probe_call
names = {0: "f32 (4 B)", 1: "v2.f32 (8 B)", 2: "v4.f32 (16 B)", 3: "v2.f16x2 (8 B)", 4: "v4.f16x2 (16 B)"}
blocks, per = 148 * 16, 64
lanes = blocks * 256 * per
for mb in (32, 48, 64, 98, 512):
    rows = mb * (1 << 20) // 16
    tab = torch.zeros(rows * 4, device="cuda")
    for mode in range(5):
        probe_call("n2m_red_bench", mode, ptr(tab), rows, blocks, per, 1, stream()); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(5):
            probe_call("n2m_red_bench", mode, ptr(tab), rows, blocks, per, 2 + r, stream())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 5
        print(f"table {mb:4d} MB  {names[mode]:16s}: {us:8.1f} us for {lanes / 1e6:.1f} M lane-REDs = {lanes / us / 1e3:7.1f} G/s, "
              f"{us * 1e-6 * 1.965e9 * 148 / lanes:5.2f} SM-cycles per lane")
    del tab
