#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native stage-0 train step (BASELINE.json metric:
ray-samples/sec of one full train step, device-timed).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm
    python bench.py --impl reference [--steps K] [--warmup W]      # CPU restatement of the reference step
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU (NCCL)

One "step" = one optimizer step of the fused pipeline on one batch of 4096 synthetic Lego-like rays
(march -> hash-grid encode -> tcgen05 MLPs -> composite + loss -> backward -> TV -> Adam), workload
"lego_stage0_converged" (SURVEY.md section 8d, config 2).  `value` = samples of all ranks / max-over-ranks
device time with the batch already resident in HBM; `e2e` = the same through Stage0Trainer.step() with
pinned-host batches (H2D inside the timed region) and a D2H read of the loss every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "lego_stage0_converged"
NUM_RAYS = 4096
ALG_BYTES = {"encode_fwd": 1053, "encode_bwd": 1024, "step": 2077}      # SURVEY.md section 8(d), bytes per sample


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same command
# (profiles/r1_ncu_summary.md, section r1f); null for kernels that were not captured
NCU_TRAFFIC = {"encode_bwd": 129.0e6 + 153.6e6, "encode_fwd": 48.78e6 + 12.57e6}


RED_LANES_PER_SAMPLE = 1.885e7 / 298645          # profiles/r1_ncu_summary.md r1f: RED sectors of k_s0_encode_bwd / samples of that launch
RED_PEAK_GLANES = 132.7                           # profiles/redbench.py: v4.f32 REDs into a 98 MB table, G lanes/s


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return float(pk["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML is polled in-process about every
    millisecond (a timed region of a few dozen sub-millisecond steps is shorter than one `nvidia-smi -lms 100` period); if NVML
    cannot be loaded the nvidia-smi loop is the fallback."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.sm, self.mx, self.reasons, self.proc = index, [], [], set(), None
        self._stop_evt = threading.Event()
        self._nv = None
        try:                                   # NVML start-up (tens of ms) happens here, before the timed region
            import pynvml as nv
            nv.nvmlInit()
            self._h = nv.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(int(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
            self._nv = nv
        except Exception:
            self._nv = None

    def _run_nvml(self):
        nv, h = self._nv, self._h
        if nv is None:
            raise RuntimeError("NVML unavailable")
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop_evt.is_set():
            self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            mask = int(get_reasons(h))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.001)

    def _run_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                      "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if r and r[0].isdigit():
                self.sm.append(int(r[0]))
            if len(r) > 1 and r[1].isdigit():
                self.mx.append(int(r[1]))
            for i in range(4):
                if len(r) >= 6 and r[2 + i].lower().startswith("active"):
                    self.reasons.add(names[i])

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def stop(self):
        self._stop_evt.set()
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


def make_batches(n_batches, seed, pin):
    from nerf2mesh_b200 import synthetic as S
    grid, bits, bricks = S.occupancy_regime("converged")
    poses = S.orbit_cameras(100, seed=0)
    intr = S.lego_intrinsics()
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        ro, rd, _, _ = S.sample_rays(poses, intr, 800, 800, NUM_RAYS, g)
        gt = S.render_bricks(ro, rd, bricks)
        bg = torch.rand(NUM_RAYS, 3, generator=g)
        noises = torch.rand(NUM_RAYS, generator=g)
        b = dict(ro=ro, rd=rd, gt=gt, bg=bg, noises=noises)
        if pin:
            b = {k: v.pin_memory() for k, v in b.items()}
        out.append(b)
    return out, grid, bits


# ------------------------------------------------------------------------------------------------
# CPU arm: the repo's PyTorch restatement of the reference step (there is no reference CPU path)
# ------------------------------------------------------------------------------------------------
def cpu_step_rate(steps, warmup, budget_s=150.0):
    """Times the CPU restatement of the train step (oracle/train_oracle.py) on a bounded sample of the 4096-ray batch.
    The thread count and the sample size are calibrated first (8-ray steps): more threads are not always faster for the
    index-heavy torch ops, and the whole (warmup + steps) run has to end within `budget_s` seconds."""
    from oracle import train_oracle as T
    torch.manual_seed(0)
    batches, grid, bits = make_batches(1, 123, False)
    cfg = dict(bound=1.0, C=1, H=128)
    ncpu = os.cpu_count() or 1

    def fresh():
        f = T.OracleField(1.0)
        return f, torch.optim.Adam(f.parameters(), lr=1e-2, eps=1e-15)

    def one(f, opt, b):
        t0 = time.perf_counter()
        _, out = T.train_step(f, opt, b["ro"], b["rd"], b["gt"], bits, cfg, b["noises"], b["bg"], "full", True)
        return time.perf_counter() - t0, out["num_points"]

    b8 = {k: v[:8] for k, v in batches[0].items()}
    best = None
    for nt in sorted({ncpu, min(ncpu, 32), min(ncpu, 8)}, reverse=True):
        torch.set_num_threads(nt)
        f, opt = fresh()
        one(f, opt, b8)                                   # first call at this setting: thread-pool start-up
        dt, _ = one(f, opt, b8)
        if best is None or dt < best[0]:
            best = (dt, nt)
    t8, threads = best
    torch.set_num_threads(threads)
    per_step = min(20.0, budget_s / max(steps + warmup, 1))
    rays = int(max(8, min(256, 8 * per_step / max(t8, 1e-3))))        # t8 / 8 over-estimates the per-ray cost (fixed overheads)
    b = {k: v[:rays] for k, v in batches[0].items()}
    f, opt = fresh()
    samples, t_total = 0, 0.0
    for it in range(warmup + steps):
        dt, m = one(f, opt, b)
        if it >= warmup:
            samples += m; t_total += dt
    return samples / t_total, t_total / max(steps, 1), threads, rays


def run_reference(args):
    """`--impl reference`: the reference has no CPU implementation of this path, so this arm times the
    CPU restatement (oracle port) with the host threads that serve it best, on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    v, sec, cores, rays = cpu_step_rate(steps, warmup)
    line = {"impl": "reference", "metric": "ray-samples/sec (train step)", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_batch": NUM_RAYS},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                             "sample": f"{rays} of the {NUM_RAYS} rays of one batch per step, full train step (oracle/train_oracle.py), "
                                       f"{warmup} warm-up + {steps} timed steps, {cores} of {os.cpu_count()} host threads (calibrated)"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (nerf2mesh_b200 has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from nerf2mesh_b200 import _lib
    from nerf2mesh_b200.stage0 import Stage0Config, Stage0Trainer
    from nerf2mesh_b200.parallel import GradSync, PeerAdam

    cfg = Stage0Config(bound=1.0, num_rays=NUM_RAYS, max_samples=NUM_RAYS * 128)
    tr = Stage0Trainer(cfg, seed=0)
    tr.nparts = args.parts
    tr.part_mode = args.part_mode
    tr.scatter_level_cuts = tuple(int(x) for x in args.scatter_cuts.split(",") if x)
    tr.level_pipe = bool(args.level_pipe)
    tr.l2_persist_mb = int(args.l2_persist_mb)
    _lib.call("n2m_s0_set_mlp_bwd_pipelined", 0 if args.mlp_bwd == "single" else 1)
    _lib.call("n2m_s0_set_mlp_bwd_issuers", 2 if args.mlp_bwd == "two-tile-2issuers" else 1)
    if args.mlp_fwd_compact:
        _lib.call("n2m_s0_set_mlp_fwd_compact", 1)
    sync = None
    dp_used = args.dp
    if world > 1:
        if args.dp == "peer":
            # all ranks must agree on the path: fall back to the NCCL all-reduce if any rank cannot map its peers
            ok = torch.ones(1, device="cuda")
            try:
                sync = PeerAdam(tr)
            except Exception as e:      # noqa: BLE001
                print(f"[rank {rank}] PeerAdam unavailable ({e}); falling back to NCCL all-reduce", file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() == 0:
                sync, dp_used = None, "nccl"
        if sync is None:
            sync = GradSync(tr)
    n_batches = 8
    host_batches, grid, bits = make_batches(n_batches, 1000 + rank, True)
    dev_batches = [{k: v.cuda(non_blocking=True) for k, v in b.items()} for b in host_batches]
    tr.set_occupancy(bits, grid)
    m_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    K, W = args.steps, args.warmup

    def tup(b):
        return (b["ro"], b["rd"], b["gt"], b["bg"], b["noises"])

    def one_step(batches, it):
        # the next batch is handed over too: its H2D copy + march overlap this step on a side stream
        b = batches[it % n_batches]
        nb = None if args.no_prefetch else tup(batches[(it + 1) % n_batches])
        tr.step(*tup(b), shading="full", use_graph=not args.no_graph, grad_sync=sync, next_batch=nb)
        m_total.add_(tr.counters[1])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- launches per step (eager, counted by the library) ----
    l0 = _lib.launch_count()
    tr.step(dev_batches[0]["ro"], dev_batches[0]["rd"], dev_batches[0]["gt"], dev_batches[0]["bg"], dev_batches[0]["noises"],
            use_graph=False, grad_sync=sync)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0

    # ---- leg 1: device-resident inputs ----
    for it in range(W):
        one_step(dev_batches, it)
    barrier()
    m_total.zero_()
    sampler = ClockSampler(local); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for it in range(K):
        one_step(dev_batches, W + it)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    samples = m_total.clone().float()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(samples, op=dist.ReduceOp.SUM)
    ms_total = ms.item(); samples_total = samples.item()
    value = samples_total / (ms_total * 1e-3)

    # ---- leg 2: end to end (pinned host -> device inside the timed region, loss read back every step) ----
    tr.drop_prefetch()
    for it in range(4):                     # untimed: (slot, parity) graph variants this leg's phase needs
        one_step(host_batches, n_batches - 4 + it)
    tr.drop_prefetch()
    barrier()
    m_total.zero_()
    loss_host = torch.zeros(4).pin_memory(); cnt_host = torch.zeros(4, dtype=torch.int32).pin_memory()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    f0.record()
    e2e_samples = 0
    for it in range(K):
        one_step(host_batches, it)
        loss_host.copy_(tr.loss_acc, non_blocking=True); cnt_host.copy_(tr.counters[:4], non_blocking=True)
        torch.cuda.current_stream().synchronize()            # the user-visible result of the step
        e2e_samples += int(cnt_host[1])
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device="cuda"); s2 = torch.tensor([float(e2e_samples)], device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX); dist.all_reduce(s2, op=dist.ReduceOp.SUM)
    e2e_value = s2.item() / (ms2.item() * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())
    d2h = loss_host.numel() * 4 + cnt_host.numel() * 4

    # ---- per-stage device times (eager, CUDA events on the launching stream) -> roofline of the dominant kernel ----
    tr.drop_prefetch()
    torch.cuda.synchronize()
    stages = ["march", "encode_fwd", "tv", "mlp_fwd", "composite_loss", "mlp_bwd", "encode_bwd", "adam"]
    acc = {s: 0.0 for s in stages}
    reps = 5
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")     # > 126 MB L2
    for r in range(reps):
        b = dev_batches[r % n_batches]
        tr.rays_o.copy_(b["ro"]); tr.rays_d.copy_(b["rd"]); tr.gt.copy_(b["gt"]); tr.bg.copy_(b["bg"]); tr.noises.copy_(b["noises"])
        tr.loss_acc.zero_()
        for s in stages:
            flush.fill_(0.0)
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); getattr(tr, s)(); z.record()
            torch.cuda.synchronize()
            acc[s] += a.elapsed_time(z) / reps
    M_last = int(tr.counters[1].item())
    peak, peak_kind = load_peaks()
    dom = max(("encode_fwd", "encode_bwd"), key=lambda s: acc[s])
    achieved = ALG_BYTES[dom] * M_last / (acc[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "k_s0_" + dom, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(dom), "alg_bytes_per_sample": ALG_BYTES[dom], "samples_per_launch": M_last,
                "kernel_ms": acc[dom], "stage_ms_cold_l2": {k: round(v, 4) for k, v in acc.items()},
                "step_frac_of_hbm": ALG_BYTES["step"] * value / 1e9 / peak}
    if dom == "encode_bwd":
        # supplementary ruler: the scatter is bound by the rate of spread REDs into a table of its footprint, not by HBM bytes
        # (profiles/redbench.py: 133 G lane-REDs/s into 98 MB, payload-independent; 63.1 lane-REDs per sample after run merging,
        # from the l1tex RED sector count of the committed ncu capture)
        lanes = RED_LANES_PER_SAMPLE * M_last
        roofline["red_rate"] = {"achieved": lanes / (acc[dom] * 1e-3) / 1e9, "peak": RED_PEAK_GLANES, "unit": "G lane-REDs/s",
                                "frac": lanes / (acc[dom] * 1e-3) / 1e9 / RED_PEAK_GLANES, "lanes_per_sample": RED_LANES_PER_SAMPLE}

    if rank == 0:
        cpu = None
        if world == 1 and not args.skip_cpu:
            v, sec, cores, rays = cpu_step_rate(1, 1, budget_s=30.0)
            cpu = {"value": v, "unit": "samples/s", "cores": cores, "kind": "port",
                   "sample": f"{rays} of the {NUM_RAYS} rays of one batch, full train step (oracle/train_oracle.py), 1 warm-up + 1 timed, "
                             f"{cores} of {os.cpu_count()} host threads (calibrated)"}
        line = {"metric": "ray-samples/sec (train step)", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "rays_per_batch": NUM_RAYS, "global_rays": NUM_RAYS * world,
                           "samples_per_step": samples_total / K, "parallelism": f"dp{world}" + ("" if world == 1 else f"-{dp_used}"), "cuda_graph": not args.no_graph, "ray_range_parts": args.parts, "part_mode": args.part_mode, "scatter_level_cuts": list(tr.scatter_level_cuts), "mlp_fwd_compact": bool(args.mlp_fwd_compact), "level_pipe": bool(args.level_pipe), "mlp_bwd": args.mlp_bwd, "l2_persist_mb": int(args.l2_persist_mb), "march_prefetch": not args.no_prefetch,
                           "l2": "inputs cycle over 8 batches; tables+grads+Adam state (0.6 GB touched per step) exceed the 126 MB L2"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms2.item() / K},
                "gpu_launches": int(launches_per_step * K), "launches_per_step": int(launches_per_step),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-prefetch", action="store_true", help="do not overlap the next batch's march with this step")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--part-mode", default="chains", choices=["pipeline", "chains"])
    ap.add_argument("--mlp-fwd-compact", action="store_true", help="experimental: MLP forward with the compact smem layout (3 CTAs/SM)")
    ap.add_argument("--mlp-bwd", default="single", choices=["single", "two-tile", "two-tile-2issuers"],
                    help="MLP backward kernel: one tile per CTA (default), two tiles + one issuer warp, or (experimental) two issuer warps")
    ap.add_argument("--l2-persist-mb", type=int, default=0,
                    help="experimental (with --scatter-cuts and --level-pipe): persisting-L2 carve-out for the gradient rows of the active scatter pass")
    ap.add_argument("--level-pipe", action="store_true",
                    help="experimental (with one --scatter-cuts level): optimizer of the first level range under the scatter of the second")
    ap.add_argument("--scatter-cuts", default="", help="experimental: comma-separated hash levels at which the scatter is cut into "
                                                       "separate launches (e.g. 10 = levels 0-9, then 10-15)")
    ap.add_argument("--parts", type=int, default=2, choices=[1, 2, 4, 8],
                    help="ray-range parts run as concurrent gather->MLP->composite->MLP'->scatter chains on forked streams")
    ap.add_argument("--dp", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = fused reduce-scatter+Adam+all-gather over NVLink peer memory, 'nccl' = all-reduce + replicated Adam")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
