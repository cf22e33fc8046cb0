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
    # after the Poisson solve: UpdatePressure + Projection (poisson.c:3137, 2700) with stand-in pressure fields
    ctx.upload("P", np.ascontiguousarray(x[..., 0])); ctx.upload("PHI", np.ascontiguousarray(0.01 * x[..., 1])); ctx.upload("UCONT", x)
    ctx.UpdatePressure()
    ctx.Projection(0.9, 0.1)
    for n in ("P", "PHI", "UCONT"):
        out["PROJ_" + n] = ctx.download(n)
    return out


def default_cases(world):
    return (("c2_box256", (61, 45, 16 * world + 7)), ("c3_turbine", (53, 37, 12 * world + 9)))


def variant_cases(world):
    """The flux / LES variants whose extra planes (Adv1-3 of the skew-symmetric form, the Clark gradient planes) live outside
    the main scalar pool and travel through the same halo layer: bitwise as well."""
    return (("c2_box256", (45, 29, 12 * world + 5), dict(skew=1, clark=1, levelset_weno=5), 0.0),
            ("c3_turbine", (37, 25, 12 * world + 3), dict(inviscid=1), 0.0))


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


def _stage(rank, msg):
    import os, sys
    if os.environ.get("VFS_SELFCHECK_TRACE"):
        print("[selfcheck rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)


def nrank_solver_and_actuators(capi, cases, rank, world, device, make_halo, verbose=True):
    """The pieces with a genuine all-reduce: vfs_momentum_solve (dot products and norms summed with ncclAllReduce),
    vfs_calc_u_lagr (element sums), plus vfs_calc_f_eul and vfs_pressure_gradient on slabs.  N ranks against 1 rank:
    same iteration counts, residual-norm history and iterate to 1e-10; F_eul and dP bitwise; U_lagr to 1e-12."""
    ok = True
    cfg = cases.scaled(cases.CONFIGS["c3_turbine"], 41, 29, 12 * world + 7)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    xyz = cases.make_grid(cfg)
    rng = np.random.default_rng(11)
    n = 24
    ii, jj, kk = rng.integers(2, mx - 3, n), rng.integers(2, my - 3, n), rng.integers(2, mz - 3, n)
    h = np.linalg.norm(xyz[2, 2, 2] - xyz[1, 1, 1])
    act = dict(cent=xyz[kk, jj, ii] + 0.4 * h * rng.uniform(-1, 1, (n, 3)), F_lagr=rng.uniform(-1, 1, (n, 3)), dA=rng.uniform(0.5, 1.5, n) * h * h,
               win=np.stack([np.maximum(ii - 3, 1), np.minimum(ii + 4, mx - 1), np.maximum(jj - 3, 1), np.minimum(jj + 4, my - 1),
                             np.maximum(kk - 3, 1), np.minimum(kk + 4, mz - 1)], -1).astype(np.int32))
    res = []
    for nr in (1, world):
        kofs, nzl = capi.slab_partition(mz, nr)[rank if nr > 1 else 0]
        p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], kofs=kofs, nzl=nzl, rank=rank if nr > 1 else 0, nranks=nr, device=device)
        ctx = capi.VfsContext(p)
        if nr > 1:
            make_halo(ctx, cfg)
        sl = slice(kofs, kofs + nzl)
        ctx.upload("COOR", xyz[sl]); ctx.FormMetrics()
        if nr == 1:
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
            for key in ("rhs_o", "dp", "f_eul"):       # a solvable step: no constant forcing on masked components (tests/solver_common.py)
                f[key] = np.zeros_like(f[key])
        for k, nm in FIELDS_IN:
            ctx.upload(nm, f[k][sl])
        ctx.upload("P", f["p"][sl])
        out = {}
        _stage(rank, "nr=%d uploads done" % nr)
        ctx.Pressure_Gradient(0.0); out["dP"] = ctx.download("DP")
        ctx.upload("DP", f["dp"][sl])
        _stage(rank, "Pressure_Gradient done")
        ctx.Contra2Cart()
        out["U_lagr"] = ctx.Calc_U_lagr([act])[0]
        _stage(rank, "Calc_U_lagr done")
        ctx.Calc_F_eul([act], df=10); out["F_eul"] = ctx.download("F_EUL")
        ctx.upload("F_EUL", f["f_eul"][sl])
        _stage(rank, "Calc_F_eul done")
        ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
        ctx.upload("UCONT", f["ucont"][sl])
        _stage(rank, "LES done, solving")
        info = ctx.momentum_solve(max_newton=3, restart=4, use_ew=0, ksp_rtol=1e-6)
        _stage(rank, "solve done: %r" % (info,))
        out["U"] = ctx.download("UCONT"); out["info"] = info
        res.append((out, sl))
        ctx.close()
    (a, _), (b, sl) = res
    checks = [("dP", np.array_equal(b["dP"], a["dP"][sl])), ("F_eul", np.array_equal(b["F_eul"], a["F_eul"][sl])),
              ("U_lagr", bool(np.abs(b["U_lagr"] - a["U_lagr"]).max() <= 1e-12 * np.abs(a["U_lagr"]).max())),
              ("solver iterations", a["info"]["ksp_its_history"] == b["info"]["ksp_its_history"] and a["info"]["reason"] == b["info"]["reason"]),
              ("solver |F| history", bool(np.abs(np.array(a["info"]["fnorm_history"]) - np.array(b["info"]["fnorm_history"])).max() <= 1e-10 * a["info"]["fnorm0"])),
              ("solver iterate", bool(np.abs(b["U"] - a["U"][sl]).max() <= 1e-10 * np.abs(a["U"]).max()))]
    for name, good in checks:
        ok = ok and good
        if not good and verbose:
            print("rank %d solver/actuator check: %s MISMATCH" % (rank, name), flush=True)
    return ok
