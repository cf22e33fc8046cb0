#!/usr/bin/env python3
"""bench.py — RHS+LES cell-updates/s (FP64) of the B200-native VFS-Wind momentum path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one cell-update pass over the whole grid: Contra2Cart + dynamic
Smagorinsky Cs + eddy viscosity + one FormFunction_SNES residual.

Workloads (BASELINE.json `configs`):
  c2_box256   synthetic stretched curvilinear box 256^3, dynamic Smagorinsky — the N = 1 default (configs[1])
  c5_weak     the weak-scaling grid 2048 x 1024 x 512: every GPU owns a 2048 x 1024 x 64 k-slab, k grows with N
              (N = 8 is configs[4] exactly) — the N > 1 default
  c3_turbine  512 x 256 x 256 turbine grid with IBM masks and F_eul (configs[2]), one GPU
  c4_farm     1024 x 512 x 256 wind farm, STRONG scaling over N k-slabs (configs[3])
  c2_weak     256 x 256 x 256N (the round-1 weak-scaling shape)
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

# algorithmic (compulsory) HBM bytes per cell-update; derivation in DESIGN.md section 5
BYTES_STEP = 248.0          # fused RHS+LES unit: 23 doubles read + 8 written
BYTES_STEP_FEUL = 272.0     # + F_eul (3 doubles)
# KERNEL-PRIVATE bytes per cell (doubles read + written by that kernel group, intermediates included) — only used for
# the per-kernel entries of `kernels`, never for `roofline`
KERNEL_PRIVATE_BYTES = {
    "c2c": 17 * 8.0,        # r: ucont3 metrics9 aj nvert1, w: ucat3
    "flux": 36 * 8.0,       # r: ucat3 nvert1 ucont3 metrics9 1/aj nu_t1, w: Fc9 Fv9
    "fp": 22 * 8.0,         # r: Fc9 Fv9 nvert1, w: Fp3   (0 when Fp is folded into the projection)
    "project": 29 * 8.0,    # r: Fp3 metrics9 1/aj nvert1 ucont3 ucont_o3 rhs_o3 dp3, w: rhs3
    "les1": 29 * 8.0,       # r: ucat3 metrics9 aj 1/aj nvert1, w: |S|1 ucat_f3 w1 U3 |S|S_ij6
    "les2": 41 * 8.0,       # r: ucat3 w1 U3 |S|S_ij6 metrics9 aj gridfactors12 ucat_f3 nvert1, w: LM MM
    "les3": 5 * 8.0,        # r: LM MM 1/aj nvert, w: Cs
    "nut": 5 * 8.0,         # r: Cs |S| aj nvert, w: nu_t
}
TIMER = {"total": 0, "c2c": 1, "flux": 2, "fp": 3, "project": 4, "les1": 5, "les2": 6, "les3": 7, "nut": 8}
METRIC = "RHS+LES cell-updates/s (FP64)"


# ---- workloads ---------------------------------------------------------------------------------------------
def resolve_workload(name, world):
    if name in (None, "auto"):
        return "c2_box256" if world == 1 else "c5_weak"
    return name


def workload_cfg(cases, name, nranks):
    """Global grid of the run and how it scales with the number of ranks."""
    if name == "c5_weak":                    # 64 node planes per GPU; 8 GPUs = the 2048 x 1024 x 512 grid of configs[4]
        cfg = dict(cases.CONFIGS["c5_weak"])
        cfg["KM"] = 64 * nranks - 1
        cfg["weak_k"] = True
        return cfg, "weak"
    if name == "c2_weak":
        cfg = dict(cases.CONFIGS["c2_box256"])
        cfg["KM"] = 256 * nranks - 1
        cfg["weak_k"] = True
        return cfg, "weak"
    cfg = dict(cases.CONFIGS[name])
    return cfg, ("strong" if nranks > 1 else "weak")


def describe(cfg, name, world, scaling):
    """The `config` object: identical in the `ours` and `reference` arms."""
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    fl = cfg["flags"]
    what = {"c2_box256": "synthetic stretched curvilinear box", "c2_weak": "synthetic stretched curvilinear box, weak scaling in k",
            "c5_weak": "weak-scaling curvilinear grid, 2048x1024x64 node k-slab per GPU (8 GPUs = 2048x1024x512)",
            "c3_turbine": "Test_09-style turbine grid with IBM nvert masks and actuator forcing F_eul",
            "c4_farm": "wind-farm ABL LES grid with turbine rows (IBM masks, F_eul), k-slab sharded"}.get(name, name)
    return {"workload": "%s: %s, %dx%dx%d nodes, dynamic Smagorinsky (les=%d), %s central%s%s" % (
                name, what, mx, my, mz, fl.get("les", 0), "2nd-order" if fl.get("second_order") else "4th-order",
                ", ii periodic" if fl.get("ii_periodic") else "", ", kk periodic" if fl.get("kk_periodic") else ""),
            "nodes": [mx, my, mz], "cells": (mx - 2) * (my - 2) * (mz - 2), "n_slabs": world, "scaling": scaling,
            "l2": "inputs larger than L2: every step streams the whole FP64 state (>= 0.3 GB per scalar field per GPU)"}


def measured_traffic(workload):
    """profiles/traffic.json: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of every kernel
    group and their per-step sum, from the committed `ncu --set full` capture of the same workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    return t if t.get("workload") == workload else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms.  Started BEFORE the warm-up (nvidia-smi needs a few hundred
    ms to come up), then only the samples whose timestamps fall inside the timed region [mark_begin, mark_end] are used."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.p, self.t0, self.t1 = dev, None, None, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        if self.t1 is None:
            self.t1 = time.time()
        time.sleep(0.05)
        self.p.terminate()
        try:
            txt = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            return out
        rows = []
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[5:9]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is None or (self.t0 - 0.01 <= r[0] <= self.t1 + 0.01)]
        note = None
        if not inside and rows:          # a timed region shorter than the sampling period: the sample nearest to it
            mid = 0.5 * ((self.t0 or self.t1) + self.t1)
            inside = [min(rows, key=lambda r: abs(r[0] - mid))]
            note = "timed region shorter than the 20 ms sampling period: nearest sample"
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if inside:
            out = {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)), "reasons": sorted(reasons), "samples": len(inside)}
            if note:
                out["note"] = note
        return out


def pin_to_gpu_numa(dev):
    """Bind this process (and the pinned host buffers it allocates afterwards: first touch) to the CPUs of the GPU's
    NUMA node, so that at N = 8 the ranks' host<->device copies do not all cross one socket."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def build_case_on_device(pkg, cfg, rank, nranks, device, halo=None, on_device=None):
    """Create the context for this rank's k-slab and fill it with the seeded synthetic state.
    Inputs are generated per slab (seed + rank) so no host ever holds the multi-GPU grid; large slabs are
    generated on the GPU itself (cases_device.py: torch, same recipe) and handed over as device pointers."""
    capi, cases = pkg.capi, pkg.cases
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    kofs, nzl = capi.slab_partition(mz, nranks)[rank]
    p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], kofs=kofs, nzl=nzl, rank=rank, nranks=nranks, device=device)
    ctx = capi.VfsContext(p)
    if halo == "nccl":
        import torch
        import torch.distributed as dist
        ctx.nccl_init(dist, device=torch.device("cuda", device))
    elif halo is not None:
        halo.attach(ctx)
    if on_device is None:
        on_device = nzl * my * mx > 40e6
    if on_device:
        f = pkg.cases_device.fill_context(ctx, cfg, kofs, nzl, rank, device)
    else:
        # grid: the slab's node planes of the global grid (coordinates depend on global indices only)
        sub = dict(cfg)
        xyz = cases.make_grid(cfg) if nranks == 1 else cases.make_grid_slab(cfg, kofs, nzl)
        ctx.upload("COOR", xyz)
        ctx.FormMetrics()
        met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
        sub["KM"] = nzl - 1
        sub["seed"] = cfg["seed"] + rank
        if nranks > 1 and cfg.get("masks"):      # bodies placed in GLOBAL coordinates, cropped to the slab
            sub["mask_window"] = (kofs, mz)
        f = cases.make_fields(sub, met)
        for k, n in (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"),
                     ("rhs_o", "RHS_O"), ("dp", "DP"), ("f_eul", "F_EUL")):      # (Ucont_rm1: BDF2 only, never read here)
            ctx.upload(n, f[k])
    return ctx, f, (mx, my, mz, kofs, nzl)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(lrank)
    ncpu_pinned = pin_to_gpu_numa(lrank) if world > 1 else 0
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    # a non-default stream: CUDA graphs cannot be captured on the legacy default stream
    torch.cuda.set_stream(torch.cuda.Stream())
    pkg = load_package()
    pkg.capi.load()
    wname = resolve_workload(args.workload, world)
    cfg, scaling = workload_cfg(pkg.cases, wname, world)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def rmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- N > 1: the N-rank result must be bitwise the 1-rank result before anything is timed ----
    parity = None
    if world > 1 and not args.no_parity:
        def make_halo(c, cf):
            c.nccl_init(dist, device=torch.device("cuda", lrank))
        ok = pkg.selfcheck.nrank_equals_1rank(pkg.capi, pkg.cases, rank, world, lrank, make_halo)
        t = torch.tensor([1.0 if ok else 0.0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        parity = "pass" if t.item() == 1.0 else "fail"
        if parity == "fail":
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": world, "multi_gpu_parity": "fail", "value": None,
                                  "error": "N-rank result differs from the 1-rank result (vfs-wind_b200/selfcheck.py); nothing was timed"}))
            dist.destroy_process_group()
            sys.exit(1)

    halo = None
    if world > 1:       # in-library NCCL k-halo layer (VFS_HALO=torch: the torch.distributed callback instead)
        halo = pkg.halo.TorchHalo(rank, world, periodic_k=bool(cfg["flags"].get("kk_periodic")), device=torch.device("cuda", lrank)) \
            if os.environ.get("VFS_HALO") == "torch" else "nccl"
    ctx, f, (mx, my, mz, kofs, nzl) = build_case_on_device(pkg, cfg, rank, world, lrank, halo)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    for key, env in ((0, "VFS_FUSED"), (2, "VFS_LES2_TY"), (3, "VFS_FLUX_MINB"), (4, "VFS_LES1_VAR"), (5, "VFS_LES3_VAR"), (6, "VFS_FASTPATH"), (7, "VFS_FLUX_VAR"),
                     (8, "VFS_FUSE_REFRESH"), (9, "VFS_OVERLAP"), (12, "VFS_FP_FUSED"), (14, "VFS_HALO_TRIM"), (15, "VFS_BOX_SHAPE"), (16, "VFS_LES_REPLAY"), (17, "VFS_FP_PAIRS"), (18, "VFS_UNIT_OVERLAP"), (19, "VFS_LES3_MINB"), (20, "VFS_BOX_OCC"), (21, "VFS_NUT_IN_LES3")):            # tuning knobs (see vfs_set_option)
        if os.environ.get(env):
            ctx.set_option(key, int(os.environ[env]))
    cells_total = (mx - 2) * (my - 2) * (mz - 2)
    k_int = [k for k in range(kofs, kofs + nzl) if 1 <= k <= mz - 2]
    cells_rank = (mx - 2) * (my - 2) * len(k_int)

    # ---- device-resident metric ("value") ----
    # the step is replayed as a CUDA graph (1st warm-up step eager, 2nd captured); the in-library NCCL halo
    # exchanges are captured with it
    use_graph = (world == 1 or halo == "nccl") and not args.no_graph
    ctx.set_option(1, 1 if use_graph else 0)
    sampler = ClockSampler(lrank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        ctx.rhs_les_fused()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.rhs_les_fused()
    e1.record(stream)
    barrier()
    sampler.mark_end()
    ms = rmax(e0.elapsed_time(e1))
    clocks = sampler.stop()

    # the two halves of the unit on their own (the Krylov solver calls the residual 10-50x per LES update)
    def timed(fn, n):
        fn(); barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream); barrier()
        return rmax(a.elapsed_time(b) / n)
    ms_rhs = timed(ctx.FormFunction_SNES_dev, args.steps)

    def les_only():
        ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
    ms_les = timed(les_only, args.steps)
    # per-kernel CUDA-event timers and the launch count come from eager steps of the same work
    ctx.set_option(1, 0)
    tsum = {k: 0.0 for k in TIMER}
    ctx.rhs_les_fused()
    l0 = ctx.launch_count()
    h0 = ctx.halo_count()            # (the graph replays above do not pass through the host-side counters)
    nt = 3
    for _ in range(nt):
        ctx.rhs_les_fused()
        for k, t in TIMER.items():
            tsum[k] += ctx.last_ms(t)
    h1 = ctx.halo_count()
    launches_step = (ctx.launch_count() - l0) // nt
    launches = launches_step * args.steps
    barrier()
    if world > 1:
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms / args.steps
    value = cells_total / (ms_step * 1e-3)

    # ---- device-resident Newton-Krylov solve: the residual is evaluated where the Krylov vectors live ----
    solver = None
    vec_bytes = nzl * my * mx * 3 * 8
    free_b = torch.cuda.mem_get_info()[0]
    if world > 1:                       # every rank must run the same number of Krylov iterations
        t = torch.tensor([float(free_b)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        free_b = float(t.item())
    fit = int(free_b * 0.9 / vec_bytes) - 9          # restart + 8 work vectors + the pinned-staging counterpart
    krylov = min(args.solver_krylov, fit)
    if hasattr(ctx, "momentum_solve") and not args.no_solver and krylov < 2:
        solver = {"skipped": "no room for a Krylov basis beside the %.0f GB of resident state (%.1f GB free, %.1f GB per vector)" % (ctx.scalar_len * 8 * ctx.nscalars / 1e9, free_b / 1e9, vec_bytes / 1e9)}
    elif hasattr(ctx, "momentum_solve") and not args.no_solver:
        try:
            args.solver_krylov = krylov
            solver = bench_solver(ctx, torch, stream, barrier, rmax, cells_total, f, (nzl, my, mx), args)
        except Exception as e:      # noqa
            solver = {"error": str(e)[:200]}

    # ---- end-to-end through the C ABI with host buffers ("e2e") ----
    e2e = None
    if not args.no_e2e:
        xh = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()
        fh = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()
        if isinstance(f.get("ucont"), np.ndarray):
            xh.copy_(torch.from_numpy(f["ucont"]))
        else:
            ctx.download_ptr("UCONT", xh.data_ptr())
        csh = torch.empty((nzl, my, mx), dtype=torch.float64).pin_memory()
        nuh = torch.empty((nzl, my, mx), dtype=torch.float64).pin_memory()
        nut_bytes = 2 * nzl * my * mx * 8
        ctx.set_option(11, 1)      # asynchronous compute-only entry points: X's upload (own stream) overlaps the LES kernels

        def e2e_step():
            ctx.upload_ptr("UCONT", xh.data_ptr())        # host lUcont -> device (what the glue does for Contra2Cart)
            ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
            # LES results back to the host Vecs: asynchronous copies on the library's copy stream, so this device->host
            # traffic overlaps the host->device copy of X below (PCIe is full duplex); waited for before the step ends
            ctx.download_async("CS", csh.data_ptr(), 0); ctx.download_async("NU_T", nuh.data_ptr(), 1)
            ctx.FormFunction_SNES(xh.data_ptr(), fh.data_ptr())   # X (host) -> F (host)
            ctx.download_wait()
            return float(fh[nzl // 2, my // 2, mx // 2, 2]) + float(nuh[nzl // 2, my // 2, mx // 2])      # read the step's results on the host
        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        n_e2e = max(1, min(args.steps, 5))
        for _ in range(n_e2e):
            e2e_step()
        e1.record(stream)
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms_e2e = rmax(max(wall, e0.elapsed_time(e1)))           # host-synchronous copies: wall clock is the honest one
        ctx.set_option(11, 0)
        e2e = {"value": cells_total / (ms_e2e / n_e2e * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": 2 * nzl * my * mx * 3 * 8,
               "d2h_bytes_per_step": nzl * my * mx * 3 * 8 + nut_bytes, "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e,
               "what": "per step: lUcont up, Contra2Cart + Cs + nu_t, Cs and nu_t down, FormFunction_SNES(X host) -> F host; pinned buffers",
               "cpus_bound_to_gpu_numa_node": ncpu_pinned}

    # ---- roofline of the whole step at the SURVEY 8(d) bytes (rank 0's share of the cells) ----
    peak, peak_src = peaks()
    per = {k: tsum[k] / nt for k in TIMER}
    bytes_step = BYTES_STEP_FEUL if cfg.get("forcing") else BYTES_STEP
    ach_step = bytes_step * cells_rank / (ms_step * 1e-3) / 1e9
    tr = measured_traffic(wname) if world == 1 else None
    step_traffic = tr.get("dram_bytes_per_step") if tr else None
    roof = {"bound": "hbm", "kernel": "whole RHS+LES step (%d launches replayed as one CUDA graph)" % launches_step, "achieved": ach_step, "peak": peak, "unit": "GB/s",
            "frac": ach_step / peak, "traffic": step_traffic, "peak_source": peak_src, "algorithmic_bytes_per_cell": bytes_step, "cells_per_launch": cells_rank,
            "wasted_traffic_ratio": (step_traffic / (bytes_step * cells_rank)) if step_traffic else None,
            "formula": "algorithmic_bytes_per_cell * cells_per_launch / ms_per_step / peak (SURVEY 8d)"}
    kernels = {}
    for k in TIMER:
        if k == "total" or per[k] <= 0:
            continue
        kernels[k] = {"ms": per[k], "kernel_private_bytes_per_cell": KERNEL_PRIVATE_BYTES[k],
                      "kernel_private_GBps": KERNEL_PRIVATE_BYTES[k] * cells_rank / (per[k] * 1e-3) / 1e9,
                      "dram_bytes_measured": (tr or {}).get("dram_bytes_per_launch", {}).get(k)}
    kernels["note"] = "kernel-private bytes count each kernel's own reads and writes (intermediates included); they are NOT the roofline figure"

    config = describe(cfg, wname, world, scaling)
    line = {"metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "run": {"k_slab_per_gpu": nzl, "resident_GB_per_gpu": ctx.scalar_len * 8 * ctx.nscalars / 1e9, "cuda_graph": bool(use_graph), "dynamic_freq": 1,
                    "inputs": "generated on the device (torch)" if not isinstance(f.get("ucont"), np.ndarray) else "generated on the host (numpy), uploaded"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": roof, "kernels": kernels,
            "rhs_only": {"value": cells_total / (ms_rhs * 1e-3), "unit": "cell-updates/s", "ms": ms_rhs, "algorithmic_bytes_per_cell": 240.0,
                         "roofline_frac": 240.0 * cells_rank / (ms_rhs * 1e-3) / 1e9 / peak, "what": "one FormFunction_SNES residual (X device resident)"},
            "les_only": {"value": cells_total / (ms_les * 1e-3), "unit": "cell-updates/s", "ms": ms_les, "algorithmic_bytes_per_cell": 128.0,
                         "roofline_frac": 128.0 * cells_rank / (ms_les * 1e-3) / 1e9 / peak, "what": "Contra2Cart + dynamic Cs + nu_t"},
            "halo": {"layer": "in-library NCCL send/recv" if halo == "nccl" else ("torch.distributed callback" if halo is not None else "single rank (periodic wrap kernels)"),
                     "exchanges_per_step": (h1[0] - h0[0]) / nt, "bytes_sent_per_step": (h1[1] - h0[1]) / nt,
                     "ghost_layers": "per exchange: only the layers the field is read at (2-4 of G = 4), see vfs_ctx.cu halo_k call sites"}}
    if e2e is not None:
        line["e2e"] = e2e
    if solver is not None:
        line["solver"] = solver
    if parity is not None:
        line["multi_gpu_parity"] = parity
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_sample(wname, planes=10, steps=2)
        try:
            line["cpu_baseline_c1"] = cpu_reference_c1()
        except Exception as e:      # noqa
            line["cpu_baseline_c1"] = {"error": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def bench_solver(ctx, torch, stream, barrier, rmax, cells_total, f, shape, args):
    """One implicit momentum solve per step through vfs_momentum_solve (device-resident GMRES + MFFD + Newton):
    Ucont uploaded from pinned host memory, solution downloaded, LES update once per step."""
    nzl, my, mx = shape
    x0 = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()
    ctx.download_ptr("UCONT", x0.data_ptr())
    out = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()

    def step():
        ctx.upload_ptr("UCONT", x0.data_ptr())
        ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
        # a fixed amount of work: `solver_newton` Newton steps of exactly `solver_krylov` GMRES iterations each
        info = ctx.momentum_solve(max_newton=args.solver_newton, max_krylov=args.solver_krylov, restart=args.solver_krylov, rtol=1e-30, atol=0.0,
                                  use_ew=0, ksp_rtol=1e-30, trust_region=0)      # (no trust-region step-length stop: the synthetic state is far from converged)
        ctx.download_ptr("UCONT", out.data_ptr())
        return info
    info = step()
    barrier()
    n = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(n):
        info = step()
    barrier()
    ms = rmax((time.perf_counter() - t0) * 1e3 / n)
    ctx.upload_ptr("UCONT", x0.data_ptr())
    ctx.momentum_release()           # the Krylov basis leaves HBM again
    nres = info["residual_evals"]
    return {"what": "per time step: lUcont up (pinned), LES update, Newton-GMRES(restart %d) with MFFD residuals on the device, Ucont down" % args.solver_krylov,
            "ms_per_time_step": ms, "residual_evals_per_step": nres, "krylov_iterations": info["krylov_iterations"], "newton_iterations": info["newton_iterations"],
            "ms_per_residual_eval_e2e": ms / max(1, nres), "rhs_cell_updates_per_s_e2e": cells_total * nres / (ms * 1e-3),
            "h2d_bytes_per_step": nzl * my * mx * 3 * 8, "d2h_bytes_per_step": nzl * my * mx * 3 * 8}


# ---- the reference's own CPU implementation (oracle/_ref), timed on the host cores -------------
def _ref_case(workload, kofs, layers, seed, plane=None):
    """Reference context for the k-slab of `layers` interior cell layers starting at global node plane `kofs` of the
    workload's grid (its own two boundary planes included; periodic k wraps the slab on itself: zero communication).
    plane = (IM, JM): the same case with a reduced i-j extent (host memory: the reference keeps ~2 KB per node)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refdrv
    import parity_common as pc
    pkg = load_package()
    cases = pkg.cases
    full = dict(cases.CONFIGS[workload])
    if plane:
        full["IM"], full["JM"] = plane
    cfg = dict(full)
    cfg["KM"] = layers + 1
    cfg["seed"] = seed
    xyz = cases.make_grid(full, kofs, layers + 2) if kofs is not None else None
    ref, xyz, f, met = pc.ref_setup(cfg, refdrv, xyz=xyz)
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    ref.view("X")[...] = f["ucont"]
    cells = (cfg["IM"] - 1) * (cfg["JM"] - 1) * (cfg["KM"] - 1)

    def step():
        ref.global_to_local("Ucont", "lUcont")
        ref.Contra2Cart(); ref.Compute_Smagorinsky_Constant_1(); ref.Compute_eddy_viscosity_LES()
        ref.FormFunction_SNES("X", "F")
    return step, cells


class quiet_stdout:
    """The reference printf()s diagnostics (wall model: "nu_t=...,ustar=...") to fd 1; bench.py's stdout is ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        nul = os.open(os.devnull, os.O_WRONLY)
        os.dup2(nul, 1)
        os.close(nul)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def _slab_worker(a):
    with quiet_stdout():
        return _slab_worker_(a)


def _slab_worker_(a):
    workload, kofs, layers, seed, warmup, steps, bar = a[:7]
    step, cells = _ref_case(workload, kofs, layers, seed, a[7] if len(a) > 7 else None)
    for _ in range(warmup):
        step()
    if bar is not None:
        bar.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return time.perf_counter() - t0, cells


def cpu_reference_sample(workload, planes, steps):
    """Bounded 1-core sample for the `cpu_baseline` of the GPU arm."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available():
        return {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libvfsref.so not present"}
    dt, cells = _slab_worker((workload if workload in ("c2_box256", "c3_turbine") else "c2_box256", 100, planes, 202, 1, steps, None))
    return {"value": cells * steps / dt, "unit": "cell-updates/s", "cores": 1, "kind": "reference",
            "sample": "reference sources (oracle/_ref) on one %d-cell-layer k-slab of the %s grid (%d cells), %d timed steps, 1 core" % (planes, workload, cells, steps)}


def run_partitioned(workload, layers_total, kofs0, cores, warmup, steps, plane=None):
    """The grid's `layers_total` interior cell layers split into `cores` contiguous k-slabs, one reference process per
    slab, all stepping concurrently; returns (seconds per step = slowest process, cells per step)."""
    import multiprocessing as mp
    cores = max(1, min(cores, layers_total // 2))
    base, rem = divmod(layers_total, cores)
    ctxm = mp.get_context("fork")
    mgr = ctxm.Manager()
    bar = mgr.Barrier(cores)
    jobs, k = [], kofs0
    for q in range(cores):
        n = base + (1 if q < rem else 0)
        jobs.append((workload, k, n, 202 + q, warmup, steps, bar, plane))
        k += n
    with ctxm.Pool(cores) as pool:
        res = pool.map(_slab_worker, jobs, chunksize=1)
    return max(r[0] for r in res) / steps, sum(r[1] for r in res), cores


def cpu_reference_c1():
    """the reference on C1 exactly (Test_10 channel on its shipped 122x42x62-node grid), 1 core and P cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available():
        return {"value": None, "sample": "oracle/_ref/libvfsref.so not present"}
    with quiet_stdout():
        step, cells = _ref_case("c1_test10", None, 60, 101)
        step()
        t0 = time.perf_counter()
        n = 3
        for _ in range(n):
            step()
        one = cells * n / (time.perf_counter() - t0)
    P = min(os.cpu_count() or 1, 64)
    sec, cellsP, used = run_partitioned("c1_test10", 60, 0, P, 1, 3)
    return {"unit": "cell-updates/s", "kind": "reference", "grid": "Test_10_ChannelFlow_Retau3000 shipped grid, 122x42x62 nodes (%d cells), wall model on" % cells,
            "value_1core": one, "value_Pcores": cellsP / sec, "cores": used, "note": "P independent k-slab processes, zero communication cost (upper bound of an MPI run)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libvfsref.so not present"}))
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    cores = min(os.cpu_count() or 1, 64)
    pkg = load_package()
    wname = resolve_workload(args.workload, world)
    cfg, scaling = workload_cfg(pkg.cases, wname, world)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    cells_total = (mx - 2) * (my - 2) * (mz - 2)
    t0 = time.perf_counter()
    # the whole grid when one pass over it takes the host cores a few seconds (c2_box256: ~2 s), otherwise a bounded
    # contiguous k-range of it (the per-cell work does not depend on the number of layers)
    base = {"c5_weak": "c5_weak", "c2_weak": "c2_box256"}.get(wname, wname)
    est_rate = 5.0e5 * cores
    budget_s = 150.0 / max(1, args.steps + args.warmup)
    layers_all = mz - 2
    # host memory: the reference keeps about 2 KB per node (its ~120 DA Vecs + work vectors), every slab process adds two
    # boundary planes — planes larger than 3e5 nodes are sampled at a reduced i-j extent (same flags, same stencils)
    plane, IMs, JMs = None, mx - 1, my - 1
    if mx * my > 3.0e5:
        fct = int(np.ceil(np.sqrt(mx * my / 3.0e5)))
        IMs, JMs = (mx - 1) // fct, (my - 1) // fct
        plane = (IMs, JMs)
    cells_plane = (IMs - 1) * (JMs - 1)
    layers = layers_all
    if plane or cells_total / est_rate > budget_s:
        layers = int(budget_s * est_rate / cells_plane)
        layers = max(2 * cores, min(layers, layers_all, int(24.0e6 / ((IMs + 1) * (JMs + 1))) - 2 * cores))
    sec, cells, used = run_partitioned(base, layers, 0, cores, args.warmup, args.steps, plane)
    value = cells / sec
    full = layers == layers_all and not plane
    sample = "reference sources (oracle/_ref): %s, %d of %d cell layers%s (%d cells per step) split into %d contiguous k-slabs, one process per core, " \
             "%d warm-up + %d timed steps, measured wall per step = slowest process; zero communication cost" % (
                 "the WHOLE grid" if full else "a bounded part of the grid", layers, layers_all,
                 " at a reduced i-j extent of %dx%d nodes" % (IMs + 1, JMs + 1) if plane else "", cells, used, args.warmup, args.steps)
    cb = {"value": value, "unit": "cell-updates/s", "cores": used, "kind": "reference", "sample": sample, "whole_grid": full}
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": describe(cfg, wname, world, scaling), "cells_per_step": cells, "cpu_baseline": cb,
            "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-solver", action="store_true")
    ap.add_argument("--solver-newton", type=int, default=1)
    ap.add_argument("--solver-krylov", type=int, default=10)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
