"""Turns the ncu artefacts of a gpurun call into the markdown tables of profiles/*_ncu_summary.md:
    python profiles/summarize_ncu.py <launches.csv> <full.ncu-rep> [<capture note>]      # with the note: also writes profiles/ncu_traffic.json
(launch list: `--metrics gpu__time_duration.sum`; full capture: `--set full`, read back with `ncu -i ... --page raw --csv`)."""
import csv, io, re, subprocess, sys
from collections import OrderedDict

def short(name):
    m = re.search(r"(k_[a-z0-9_]+)", name)
    if m:
        return m.group(1)
    m = re.search(r"at::native::([a-zA-Z_]+)|at::([a-zA-Z_]+)", name)
    return "torch:" + (m.group(1) or m.group(2)) if m else name[:40]

def launches(path):
    rows = [l for l in open(path) if l.startswith('"')]
    r = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    for x in r:
        if x["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(x["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += float(x["Metric Value"]) / 1e3
    return agg

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    want = OrderedDict([
        ("dur us", "gpu__time_duration.sum"), ("DRAM rd MB", "dram__bytes_read.sum"), ("DRAM wr MB", "dram__bytes_write.sum"),
        ("DRAM %", "dram__throughput.avg.pct_of_peak_sustained_elapsed"), ("L1 %", "l1tex__throughput.avg.pct_of_peak_sustained_active"),
        ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"), ("occ %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active"), ("tensor %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("warp inst", "smsp__inst_executed.sum"),
        ("RED sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum"), ("LD sectors", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
        ("L2 sectors", "lts__t_sectors.sum")])
    idx = {}
    for label, metric in want.items():
        exact = [i for i, h in enumerate(hdr) if h == metric]
        idx[label] = exact + [i for i, h in enumerate(hdr) if h.endswith("." + metric)]
    units = rows[1]
    res = OrderedDict()
    for row in rows[2:]:
        if len(row) < len(hdr):
            continue
        k = short(row[hdr.index("Kernel Name")])
        if k in res:
            continue                       # first captured launch of each kernel
        d = OrderedDict()
        for label, cand in idx.items():
            v = None
            for i in cand:
                try:
                    v = float(row[i].replace(",", "")); break
                except ValueError:
                    continue
            if v is None:
                continue
            u = units[i]
            if label == "dur us":
                v = v / 1e3 if u in ("ns", "nsecond") else v * {"us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
            if label.startswith("DRAM ") and label.endswith("MB"):
                v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
            d[label] = v
        res[k] = d
    return res

if __name__ == "__main__":
    agg = launches(sys.argv[1])
    ours = {k: v for k, v in agg.items() if k.startswith("k_")}
    steps = max(1, ours.get("k_adam_post", ours.get("k_dp_post", [1]))[0])
    tot = sum(v[1] / v[0] * (v[0] / steps) for v in ours.values())
    print("| kernel | launches | mean us | per step us | share of step |\n|---|---|---|---|---|")
    for k, (n, t) in ours.items():
        per_step = t / steps
        print(f"| `{k}` | {n} | {t / n:.1f} | {per_step:.1f} | {100 * per_step / tot:.1f} % |")
    print(f"| **sum** | | | **{tot:.1f}** | |\n")
    if len(sys.argv) > 2:
        res = full(sys.argv[2])
        keys = list(res)
        labels = list(next(iter(res.values())).keys())
        print("| metric | " + " | ".join(f"`{k}`" for k in keys) + " |\n|---|" + "---|" * len(keys))
        for lab in labels:
            print(f"| {lab} | " + " | ".join(f"{res[k].get(lab, float('nan')):.4g}" for k in keys) + " |")
        if len(sys.argv) > 3:
            # profiles/ncu_traffic.json: DRAM bytes per launch of every captured kernel, read by bench.py (`roofline.traffic`)
            import json, os
            out = {"capture": os.path.basename(sys.argv[2]) + " (" + sys.argv[3] + ")",
                   "kernels": {k: {"dram_bytes": (v.get("DRAM rd MB", 0.0) + v.get("DRAM wr MB", 0.0)) * 1e6, "dur_us": v.get("dur us")} for k, v in res.items()}}
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json"), "w") as f:
                json.dump(out, f, indent=1)
