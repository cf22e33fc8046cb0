#!/usr/bin/env python3
"""bench.py — RHS+LES cell-updates/s (FP64) of the B200-native VFS-Wind momentum path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one cell-update pass over the whole grid: Contra2Cart + dynamic
Smagorinsky Cs + eddy viscosity + one FormFunction_SNES residual.  N=1 workload: config[1] of
BASELINE.json (synthetic stretched curvilinear box 256^3, dynamic Smagorinsky).  N>1: weak scaling,
every rank owns a 256x256 x (256 planes) k-slab of a 256 x 256 x 256N grid, k-halos over NCCL.
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

# algorithmic (compulsory) HBM bytes per cell; derivation in DESIGN.md section 5
BYTES_STEP = 248.0          # fused RHS+LES unit, SURVEY 8(d): 23 doubles read + 8 written
BYTES_STEP_FEUL = 272.0
KERNEL_BYTES = {            # per-kernel-group compulsory traffic (doubles read + written per cell) * 8
    "flux": 36 * 8.0,       # r: ucat3 nvert1 ucont3 metrics9 1/aj nu_t1, w: Fc9 Fv9
    "fp": 22 * 8.0,         # r: Fc9 Fv9 nvert1, w: Fp3
    "project": 29 * 8.0,    # r: Fp3 metrics9 1/aj nvert1 ucont3 ucont_o3 rhs_o3 dp3, w: rhs3
    "c2c": 17 * 8.0,        # r: ucont3 metrics9 aj nvert1, w: ucat3
    "les1": 29 * 8.0,       # r: ucat3 metrics9 aj 1/aj nvert1, w: |S|1 ucat_f3 w1 U3 |S|S_ij6
    "les2": 41 * 8.0,       # r: ucat3 w1 U3 |S|S_ij6 metrics9 aj gridfactors12 ucat_f3 nvert1, w: LM MM
    "les3": 5 * 8.0,        # r: LM MM 1/aj nvert, w: Cs
    "nut": 5 * 8.0,         # r: Cs |S| aj nvert, w: nu_t
}
# kernel that dominates each timer group (names as they appear in the ncu launch list / profiles/)
KERNEL_NAME = {"flux": "k_tile_march<RingFlux, FluxBody>", "fp": "k_box<FpCell>", "project": "k_box<ProjectSNES>", "c2c": "k_box<C2CInterior>",
               "les1": "k_tile_march<RingLes1, Les1Body>", "les2": "k_les2_march<Les2MarchT<12>>", "les3": "k_filter_march<Les3March, 2>", "nut": "k_box<NuT>"}
TIMER = {"total": 0, "c2c": 1, "flux": 2, "fp": 3, "project": 4, "les1": 5, "les2": 6, "les3": 7, "nut": 8}


def measured_traffic(kernel_group, workload):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the group's dominant kernel, from the
    committed `ncu --set full` capture summarised in profiles/traffic.json (same workload, one launch)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p))
    if t.get("workload") != workload:
        return None
    return t.get("dram_bytes_per_launch", {}).get(kernel_group)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, dev):
        self.dev, self.p = dev, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.p:
            return out
        self.p.terminate()
        try:
            txt = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            return out
        sm, mx, reasons = [], [], set()
        for line in txt.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def workload_cfg(cases, name, nranks):
    cfg = dict(cases.CONFIGS[name])
    if nranks > 1:                       # weak scaling: fixed 256-plane slab per GPU
        cfg["KM"] = (cfg["KM"] + 1) * nranks - 1
        cfg["weak_k"] = True          # keep the cell size: domain length in z grows with N
    return cfg


def build_case_on_device(pkg, cfg, rank, nranks, device, halo=None):
    """Create the context for this rank's k-slab and fill it with the seeded synthetic state.
    Inputs are generated per slab (seed + rank) so no host ever holds the multi-GPU grid."""
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
    # grid: the slab's node planes of the global grid (coordinates depend on global indices only)
    sub = dict(cfg)
    xyz = cases.make_grid(cfg) if nranks == 1 else cases.make_grid_slab(cfg, kofs, nzl)
    ctx.upload("COOR", xyz)
    ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    sub["KM"] = nzl - 1
    sub["seed"] = cfg["seed"] + rank
    f = cases.make_fields(sub, met)
    if nranks > 1:
        f["nvert"][...] = 0.0 if not cfg.get("masks") else f["nvert"]
    for k, n in (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"),
                 ("ucont_rm1", "UCONT_RM1"), ("rhs_o", "RHS_O"), ("dp", "DP"), ("f_eul", "F_EUL")):
        ctx.upload(n, f[k])
    return ctx, f, (mx, my, mz, kofs, nzl)


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lrank = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    # a non-default stream: CUDA graphs cannot be captured on the legacy default stream
    torch.cuda.set_stream(torch.cuda.Stream())
    pkg = load_package()
    pkg.capi.load()
    cfg = workload_cfg(pkg.cases, args.workload, world)
    halo = None
    if world > 1:       # in-library NCCL k-halo layer (VFS_HALO=torch: the torch.distributed callback instead)
        halo = pkg.halo.TorchHalo(rank, world, periodic_k=bool(cfg["flags"].get("kk_periodic")), device=torch.device("cuda", lrank)) \
            if os.environ.get("VFS_HALO") == "torch" else "nccl"
    ctx, f, (mx, my, mz, kofs, nzl) = build_case_on_device(pkg, cfg, rank, world, lrank, halo)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    for key, env in ((0, "VFS_FUSED"), (2, "VFS_LES2_TY"), (3, "VFS_FLUX_MINB"), (4, "VFS_LES1_VAR"), (5, "VFS_LES3_VAR"), (6, "VFS_FASTPATH"), (7, "VFS_FLUX_VAR"), (8, "VFS_FUSE_REFRESH"), (9, "VFS_OVERLAP")):            # tuning knobs (see vfs_set_option)
        if os.environ.get(env):
            ctx.set_option(key, int(os.environ[env]))
    cells_total = (mx - 2) * (my - 2) * (mz - 2)
    k_int = [k for k in range(kofs, kofs + nzl) if 1 <= k <= mz - 2]
    cells_rank = (mx - 2) * (my - 2) * len(k_int)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident metric ("value") ----
    # the ~100-launch step is replayed as a CUDA graph (1st warm-up step eager, 2nd captured); the
    # in-library NCCL halo exchanges are captured with it
    use_graph = (world == 1 or halo == "nccl") and not args.no_graph
    ctx.set_option(1, 1 if use_graph else 0)
    for _ in range(max(args.warmup, 3)):
        ctx.rhs_les_fused()
    barrier()
    sampler = ClockSampler(lrank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.rhs_les_fused()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # the two halves of the unit on their own (the Krylov solver calls the residual 10-50x per LES update)
    def timed(fn, n):
        fn(); barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream); barrier()
        t = a.elapsed_time(b) / n
        if world > 1:
            tt = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t
    ms_rhs = timed(ctx.FormFunction_SNES_dev, args.steps)

    def les_only():
        ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
    ms_les = timed(les_only, args.steps)
    # per-kernel CUDA-event timers and the launch count come from eager steps of the same work
    ctx.set_option(1, 0)
    tsum = {k: 0.0 for k in TIMER}
    ctx.rhs_les_fused()
    l0 = ctx.launch_count()
    nt = 3
    for _ in range(nt):
        ctx.rhs_les_fused()
        for k, t in TIMER.items():
            tsum[k] += ctx.last_ms(t)
    launches = (ctx.launch_count() - l0) // nt * args.steps
    barrier()
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms / args.steps
    value = cells_total / (ms_step * 1e-3)

    # ---- end-to-end through the C ABI with host buffers ("e2e") ----
    xh = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()
    fh = torch.empty((nzl, my, mx, 3), dtype=torch.float64).pin_memory()
    xh.copy_(torch.from_numpy(f["ucont"]))
    xn, fn = xh.numpy(), fh.numpy()
    nut_bytes = 2 * nzl * my * mx * 8

    csh = torch.empty((nzl, my, mx), dtype=torch.float64).pin_memory()
    nuh = torch.empty((nzl, my, mx), dtype=torch.float64).pin_memory()

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
    ms_e2e = max(wall, e0.elapsed_time(e1))           # host-synchronous copies: wall clock is the honest one
    if world > 1:
        t = torch.tensor([ms_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = cells_total / (ms_e2e / n_e2e * 1e-3)
    h2d = 2 * nzl * my * mx * 3 * 8
    d2h = nzl * my * mx * 3 * 8 + nut_bytes

    # ---- roofline of the dominant kernel group (rank 0's timers) ----
    peak, peak_src = peaks()
    per = {k: tsum[k] / nt for k in TIMER}
    dom = max((k for k in TIMER if k != "total"), key=lambda k: per[k])
    ach = KERNEL_BYTES[dom] * cells_rank / (per[dom] * 1e-3) / 1e9 if per[dom] > 0 else 0.0
    bytes_step = BYTES_STEP_FEUL if cfg.get("forcing") else BYTES_STEP
    ach_step = bytes_step * value / world / 1e9
    roof = {"bound": "hbm", "kernel": dom, "kernel_name": KERNEL_NAME[dom], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": measured_traffic(dom, args.workload) if world == 1 else None,
            "peak_source": peak_src, "algorithmic_bytes_per_cell": KERNEL_BYTES[dom], "ms_per_launch_group": per[dom]}
    roof_step = {"bound": "hbm", "achieved": ach_step, "peak": peak, "unit": "GB/s", "frac": ach_step / peak, "algorithmic_bytes_per_cell": bytes_step,
                 "note": "whole fused RHS+LES step at SURVEY 8(d) bytes, per GPU"}

    line = {"metric": "RHS+LES cell-updates/s (FP64)", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: synthetic stretched curvilinear box %dx%dx%d nodes, dynamic Smagorinsky (les=2), 4th-order central, ii+kk periodic" % (args.workload, mx, my, mz),
                       "cells": cells_total, "k_slab_per_gpu": nzl, "l2": "inputs larger than L2 (%.1f GB resident state per GPU; every kernel streams >= 0.5 GB)" % (ctx.scalar_len * 8 * ctx.nscalars / 1e9),
                       "dynamic_freq": 1, "cuda_graph": bool(use_graph)},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e},
            "roofline": roof, "roofline_step": roof_step,
            "rhs_only": {"value": cells_total / (ms_rhs * 1e-3), "unit": "cell-updates/s", "ms": ms_rhs, "algorithmic_bytes_per_cell": 240.0,
                         "roofline_frac": 240.0 * cells_total / world / (ms_rhs * 1e-3) / 1e9 / peak, "what": "one FormFunction_SNES residual (X device resident)"},
            "les_only": {"value": cells_total / (ms_les * 1e-3), "unit": "cell-updates/s", "ms": ms_les, "algorithmic_bytes_per_cell": 128.0,
                         "roofline_frac": 128.0 * cells_total / world / (ms_les * 1e-3) / 1e9 / peak, "what": "Contra2Cart + dynamic Cs + nu_t"},
            "halo": {"layer": "in-library NCCL send/recv" if halo == "nccl" else ("torch.distributed callback" if halo is not None else "single rank (periodic wrap kernels)"),
                     "exchanges": ctx.halo_count()[0], "bytes_sent": ctx.halo_count()[1]},
            "kernel_ms": per}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args.workload, cores=1, steps=2, planes=10)
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# ---- the reference's own CPU implementation (oracle/_ref), timed on the host cores -------------
def _ref_worker(a):
    workload, planes, steps, seed = a
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refdrv
    import parity_common as pc
    pkg = load_package()
    cfg = dict(pkg.cases.CONFIGS[workload])
    cfg["KM"] = planes + 1            # k-slab crop: planes interior cell layers, same i-j extent
    cfg["seed"] = seed
    ref, xyz, f, met = pc.ref_setup(cfg, refdrv)
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    ref.view("X")[...] = f["ucont"]
    cells = (cfg["IM"] - 1) * (cfg["JM"] - 1) * (cfg["KM"] - 1)

    def step():
        ref.global_to_local("Ucont", "lUcont")
        ref.Contra2Cart(); ref.Compute_Smagorinsky_Constant_1(); ref.Compute_eddy_viscosity_LES()
        ref.FormFunction_SNES("X", "F")
    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return cells * steps / dt, cells


def cpu_reference(workload, cores, steps, planes):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrv
    if not refdrv.available():
        return {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libvfsref.so not present"}
    if cores == 1:
        rate, cells = _ref_worker((workload, planes, steps, 202))
        rates = [rate]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_ref_worker, [(workload, planes, steps, 202 + q) for q in range(cores)])
        rates = [r for r, _ in res]; cells = res[0][1]
    return {"value": float(sum(rates)), "unit": "cell-updates/s", "cores": cores, "kind": "reference",
            "sample": "reference sources (oracle/_ref) on %d independent %d-cell-layer k-slab crop(s) of the %s grid (%d cells each), %d timed steps each, zero communication cost" % (cores, planes, workload, cells, steps)}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    t0 = time.perf_counter()
    cb = cpu_reference(args.workload, cores=cores, steps=max(1, min(args.steps, 3)), planes=6)
    if cb["value"] is None:
        print(json.dumps({"impl": "reference", "unavailable": cb["sample"]}))
        return
    world = int(os.environ.get("WORLD_SIZE", 1))
    pkg = load_package()
    cfg = workload_cfg(pkg.cases, args.workload, world)
    cells_total = (cfg["IM"] - 1) * (cfg["JM"] - 1) * (cfg["KM"] - 1)
    ms_equiv = cells_total / cb["value"] * 1e3        # one pass over the arm's whole grid at the sampled host rate
    line = {"impl": "reference", "metric": "RHS+LES cell-updates/s (FP64)", "value": cb["value"], "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_equiv, "ms_per_step_note": "whole %d-cell grid at the rate measured on the bounded sample" % cells_total, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + " (bounded k-slab sample, see cpu_baseline.sample)"}, "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_box256")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
