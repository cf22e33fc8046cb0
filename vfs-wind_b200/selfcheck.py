"""N-rank k-slab run == 1-rank run, bitwise (the multi-GPU contract of the halo layer that replaces
DAGlobalToLocal / DALocalToLocal between ranks, Source/init.c:131-160).

Product-side self check: it compares the library with itself (N ranks against 1 rank on the same
device), never with the oracle.  Used by bench.py before it times an N > 1 run
(`"multi_gpu_parity"` in the JSON line) and by tests/multigpu_check.py.
"""
import numpy as np

FIELDS_IN = (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"),
             ("ucont_rm1", "UCONT_RM1"), ("rhs_o", "RHS_O"), ("dp", "DP"), ("f_eul", "F_EUL"))


def run_path(ctx, x):
    """The whole path through its separate entry points and through the fused unit."""
    out = {}
    ctx.Contra2Cart()
    ctx.Compute_Smagorinsky_Constant_1()
    ctx.Compute_eddy_viscosity_LES()
    out["F"] = ctx.FormFunction_SNES(x)
    for n in ("UCAT", "CS", "NU_T", "UCONT", "CSI", "AJ"):
        out[n] = ctx.download(n)
    ctx.upload("UCONT", x)
    ctx.rhs_les_fused()
    for n in ("RHS", "UCAT", "CS", "NU_T"):
        out["FUSED_" + n] = ctx.download(n)
    return out


def default_cases(world):
    return (("c2_box256", (61, 45, 16 * world + 7)), ("c3_turbine", (53, 37, 12 * world + 9)))


def homogeneous_cases(world):
    """Channel-flow setting (LM, MM averaged over i and k, les.c:798-838) and k alone: the plane / line sums are the one
    all-reduce of the path, so N ranks equal 1 rank to rounding only — (config, dims, extra flags, tolerance)."""
    return (("c2_box256", (45, 29, 12 * world + 5), dict(i_homo_filter=1, k_homo_filter=1), 1e-12),
            ("c3_turbine", (37, 25, 12 * world + 3), dict(k_homo_filter=1), 1e-12))


def nrank_equals_1rank(capi, cases, rank, world, device, make_halo, case_list=None, verbose=True):
    """Every rank computes the single-rank result on its own device, then its slab of the N-rank
    run; returns True when all compared fields of this rank are bitwise equal.  `make_halo(ctx, cfg)`
    attaches the halo layer (vfs_nccl_init or a callback) to the slab context."""
    ok = True
    for case in (case_list or default_cases(world)):
        cfgname, dims = case[0], case[1]
        extra, tol = (case[2], case[3]) if len(case) > 2 else ({}, 0.0)
        cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
        cfg["flags"] = dict(cfg["flags"], **extra)
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        xyz = cases.make_grid(cfg)
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], device=device))
        ctx.upload("COOR", xyz); ctx.FormMetrics()
        met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
        f = cases.make_fields(cfg, met)
        for k, n in FIELDS_IN:
            ctx.upload(n, f[k])
        x = f["ucont"] * (1.0 + 1e-3 * np.sin(np.arange(f["ucont"].size).reshape(f["ucont"].shape)))
        single = run_path(ctx, x)
        ctx.close()
        kofs, nzl = capi.slab_partition(mz, world)[rank]
        p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], kofs=kofs, nzl=nzl, rank=rank, nranks=world, device=device)
        ctx = capi.VfsContext(p)
        make_halo(ctx, cfg)
        sl = slice(kofs, kofs + nzl)
        ctx.upload("COOR", xyz[sl]); ctx.FormMetrics()
        for k, n in FIELDS_IN:
            ctx.upload(n, f[k][sl])
        out = run_path(ctx, x[sl])
        for n in sorted(out):
            a, b = out[n], single[n][sl]
            same = np.array_equal(a, b) if tol == 0 else bool(np.abs(a - b).max() <= tol * max(np.abs(single[n]).max(), 1e-300))
            ok = ok and same
            if not same and verbose:
                print("rank %d %s %s MISMATCH max %.3e" % (rank, cfgname, n, np.abs(out[n] - single[n][sl]).max()), flush=True)
        ctx.close()
    return ok
