"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck), e.g.
    compute-sanitizer --tool racecheck python tests/sanitize_small.py
Exercises every marching kernel of the default path on several tiles and k-chunks, then (second half) the one-thread-per-
face / per-cell variants (WENO3, skew, Clark, bctype 11), the homogeneous Cs averaging, Pressure_Gradient, UpdatePressure
+ Projection, the actuator kernels and a short device-resident solve."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import parity_common as pc
pkg = pc.load_package()
capi, cases = pkg.capi, pkg.cases


def setup(cfg):
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
    xyz = cases.make_grid(cfg)
    if cfg.get("z_shift"):
        xyz[..., 2] -= cfg["z_shift"]
    ctx.upload("COOR", xyz); ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    f = cases.make_fields(cfg, met)
    for k, n in pc.FIELDS_IN:
        ctx.upload(n, f[k])
    return ctx, xyz, f


for name, dims in (("c2_box256", (70, 37, 40)), ("c3_turbine", (45, 30, 35))):
    cfg = cases.scaled(cases.CONFIGS[name], *dims)
    ctx, xyz, f = setup(cfg)
    ctx.rhs_les_fused()
    ctx.Convection(); ctx.Viscous()
    print(name, float(np.abs(ctx.download("RHS")).max()), flush=True)
    ctx.close()

if "--default-only" not in sys.argv:
    variants = [("c2_box256", dict(skew=1, clark=1, levelset_weno=5), None, 0.0),
                ("c3_turbine", dict(inviscid=1, i_homo_filter=1, k_homo_filter=1), None, 0.0),
                ("c3_turbine", dict(ii_periodic=0, kk_periodic=0, skew=1), [11, 1, 1, 1, 5, 4], 1.4)]
    for name, extra, bctype, zshift in variants:
        cfg = cases.scaled(cases.CONFIGS[name], 37, 25, 29)
        cfg["flags"] = dict(cfg["flags"], **extra)
        if bctype:
            cfg["bctype"] = bctype
        cfg["z_shift"] = zshift
        ctx, xyz, f = setup(cfg)
        ctx.rhs_les_fused()
        ctx.upload("P", f["p"]); ctx.Pressure_Gradient(0.1); ctx.upload("DP", f["dp"])
        ctx.upload("PHI", 0.05 * f["p"]); ctx.UpdatePressure(); ctx.Projection(0.9, 0.1); ctx.Contra2Cart()
        cyl = ctx.cylinder_forces()
        act = pc.make_actuator(cfg, xyz)
        ul = ctx.Calc_U_lagr([act]); ctx.Calc_F_eul([act], df=10)
        for key in ("rhs_o", "dp", "f_eul"):
            ctx.upload(dict(rhs_o="RHS_O", dp="DP", f_eul="F_EUL")[key], np.zeros_like(f[key]))
        ctx.upload("UCONT", f["ucont"])
        info = ctx.momentum_solve(max_newton=1, restart=3, use_ew=0, ksp_rtol=1e-3)
        print(name, extra, float(np.abs(ctx.download("RHS")).max()), float(np.abs(cyl).max()), float(np.abs(ul[0]).max()), info["ksp_its_history"], flush=True)
        ctx.close()
