"""Shared parity driver: runs the same seeded case through the oracle (oracle/_ref = the
reference's own sources) and through a VfsContext, returning per-field relative errors
max|a-b| / max|b| (SURVEY 8c tolerance: 1e-12 for Rhs, Ucat, Cs, nu_t; masks bit-exact)."""
import importlib.util
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def load_package():
    name = "vfs_wind_b200"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "vfs-wind_b200", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(ROOT, "vfs-wind_b200")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def relerr(a, b):
    """max|a-b| / max|b| over the finite entries; non-finite entries (the reference itself divides by zero in
    degenerate set-ups, e.g. the wall model at a zeroed corner cell) must coincide exactly, else inf."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    fa, fb = np.isfinite(a), np.isfinite(b)
    if not (fa.all() and fb.all()):
        if not np.array_equal(fa, fb) or not np.array_equal(np.isnan(a), np.isnan(b)) or not np.array_equal(a[~fa & ~np.isnan(a)], b[~fb & ~np.isnan(b)]):
            return float("inf")
        a, b = a[fa], b[fb]
        if a.size == 0:
            return 0.0
    den = np.abs(b).max()
    if den == 0:
        return float(np.abs(a).max())
    return float(np.abs(a - b).max() / den)


def krylov_x(ucont):
    """A perturbed state like the U + h*v vectors the matrix-free Krylov solver evaluates."""
    return ucont + 1e-4 * np.abs(ucont).max() * np.cos(np.arange(ucont.size).reshape(ucont.shape) * 0.37)


def conv_defined(cfg, fields):
    """Cells where the reference's Convection output is defined: the interior, minus rhs.c:903,977 — at the
    last face of a non-periodic i/j direction whose lower neighbour is an IB/solid node the reference reads
    ucat[..][m], past the end of its array.  Shape (mz, my, mx, 1) of 0/1."""
    nv, fl = fields["nvert"], cfg["flags"]
    ok = np.zeros(nv.shape, bool)
    ok[1:-1, 1:-1, 1:-1] = True
    if not fl.get("ii_periodic"):
        ok[:, :, -2] &= ~(nv[:, :, -3] > 0.1)
    if not fl.get("jj_periodic"):
        ok[:, -2, :] &= ~(nv[:, -3, :] > 0.1)
    return ok[..., None].astype(float)


def make_actuator(cfg, xyz, seed=5, n=40, half=3):
    """Synthetic actuator elements near random interior nodes: centres offset from the node, index windows of
    +-`half` cells clipped to the interior (as the reference's pre-processing leaves them), random forces and areas."""
    rng = np.random.default_rng(seed)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ii = rng.integers(2, mx - 3, n); jj = rng.integers(2, my - 3, n); kk = rng.integers(2, mz - 3, n)
    h = np.linalg.norm(xyz[2, 2, 2] - xyz[1, 1, 1])
    cent = xyz[kk, jj, ii] + 0.4 * h * rng.uniform(-1, 1, (n, 3))
    win = np.stack([np.maximum(ii - half, 1), np.minimum(ii + half + 1, mx - 1), np.maximum(jj - half, 1), np.minimum(jj + half + 1, my - 1),
                    np.maximum(kk - half, 1), np.minimum(kk + half + 1, mz - 1)], -1).astype(np.int32)
    return dict(cent=cent, F_lagr=rng.uniform(-1, 1, (n, 3)), dA=rng.uniform(0.5, 1.5, n) * h * h, win=win)


def ref_setup(cfg, refdrv, xyz=None):
    """Create the reference context, metrics and input state for cfg (node coordinates `xyz`, default: cfg's own
    grid).  Returns (ref, xyz, fields, metrics)."""
    pkg = load_package()
    cases = pkg.cases
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ref = refdrv.RefCase(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"])
    if xyz is None:
        xyz = cases.make_grid(cfg)
        if cfg.get("z_shift"):      # e.g. the body-fitted cylinder rule of bctype 11 keys on the sign of z (rhs.c:627)
            xyz[..., 2] -= cfg["z_shift"]
    ref.set_coords(xyz)
    ref.FormMetrics()
    met = dict(csi=np.array(ref.owned("lCsi")), eta=np.array(ref.owned("lEta")), zet=np.array(ref.owned("lZet")), aj=np.array(ref.owned("lAj")))
    fields = cases.make_fields(cfg, met)
    ref.set_owned("Nvert", fields["nvert"])
    ref.global_to_local("Nvert", "lNvert")
    ref.set_owned("Ucont", fields["ucont"])
    ref.global_to_local("Ucont", "lUcont")
    ref.set_owned("Ucat", fields["ucat"])
    ref.global_to_local("Ucat", "lUcat")
    ref.set_owned("lUcat_old", fields["ucat_old"])
    ref.wrap_fill("lUcat_old")
    ref.set_owned("Ucont_o", fields["ucont_o"])
    ref.set_owned("Ucont_rm1", fields["ucont_rm1"])
    ref.set_owned("RHS_o", fields["rhs_o"])
    ref.set_owned("dP", fields["dp"])
    ref.set_owned("F_eul", fields["f_eul"])
    return ref, xyz, fields, met


def dev_setup(cfg, xyz, fields, lib=None, device=0, options=None):
    pkg = load_package()
    capi = pkg.capi
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], device=device)
    ctx = capi.VfsContext(p, lib=lib)
    for key, val in (options or {}).items():
        ctx.set_option(key, val)
    ctx.upload("COOR", xyz)
    ctx.FormMetrics()
    ctx.upload("NVERT", fields["nvert"])
    ctx.upload("UCONT", fields["ucont"])
    ctx.upload("UCAT", fields["ucat"])
    ctx.upload("UCAT_OLD", fields["ucat_old"])
    ctx.upload("UCONT_O", fields["ucont_o"])
    ctx.upload("UCONT_RM1", fields["ucont_rm1"])
    ctx.upload("RHS_O", fields["rhs_o"])
    ctx.upload("DP", fields["dp"])
    ctx.upload("F_EUL", fields["f_eul"])
    return ctx


def run_parity(cfg, refdrv, lib=None, device=0, verbose=False, options=None, legacy=True):
    """Full path comparison.  Returns dict name -> relative error."""
    ref, xyz, fields, met = ref_setup(cfg, refdrv)
    ctx = dev_setup(cfg, xyz, fields, lib=lib, device=device, options=options)
    err = {}
    for nm, key in (("CSI", "csi"), ("ETA", "eta"), ("ZET", "zet"), ("AJ", "aj")):
        err["metrics_" + nm] = relerr(ctx.download(nm), met[key])
    # Flow_Solver LES block: Contra2Cart, Cs, nu_t (solvers.c:365-371)
    ref.Contra2Cart()
    ctx.Contra2Cart()
    err["Contra2Cart_ucat"] = relerr(ctx.download("UCAT"), ref.owned("Ucat"))
    if any(b in (-1, -2) for b in cfg["bctype"][:4]):      # wall-function sides: friction velocity of the first cells (rhs.c:336,371,401,435)
        err["Contra2Cart_ustar"] = relerr(ctx.download("USTAR")[1:-1, 1:-1, 1:-1], np.array(ref.owned("lUstar"))[1:-1, 1:-1, 1:-1])
    if cfg["flags"].get("les"):
        ref.Compute_Smagorinsky_Constant_1()
        ctx.Compute_Smagorinsky_Constant_1()
        err["Cs"] = relerr(ctx.download("CS"), ref.owned("lCs"))
        ref.Compute_eddy_viscosity_LES()
        ctx.Compute_eddy_viscosity_LES()
        err["nu_t"] = relerr(ctx.download("NU_T"), ref.owned("lNu_t"))
    # legacy explicit-solver terms on the same state (rhs.c:751, 1071).  Conv is only defined
    # on the interior cells (the reference leaves the rest of its output Vec untouched).
    if legacy:
        ref.new_vec("Conv", 3, False)
        ref.new_vec("Visc", 3, False)
        ref.Convection("Conv")
        ref.Viscous("Visc")
        ctx.Convection()
        ctx.Viscous()
        ok = conv_defined(cfg, fields)
        err["Convection"] = relerr(ctx.download("CONV") * ok, np.array(ref.view("Conv")) * ok)
        err["Viscous"] = relerr(ctx.download("VISC"), ref.view("Visc"))
    # RHS_o = Formfunction_2(U) with scale 1 (solvers.c:628-629)
    ref.IB_BC()
    ctx.IB_BC()
    err["IB_BC_ucont"] = relerr(ctx.download("UCONT"), ref.owned("lUcont"))
    err["IB_BC_nvert_mismatches"] = float(np.count_nonzero(ctx.download("NVERT") != np.array(ref.owned("lNvert"))))     # masks: bit-exact
    ref.view("RHS_o")[...] = 0
    ref.Formfunction_2("RHS_o", 1.0)
    ctx.upload("RHS_O", np.zeros_like(fields["rhs_o"]))
    ctx.Formfunction_2("RHS_O", 1.0)
    rhs_o_ref = np.array(ref.owned("RHS_o"))
    err["Formfunction_2"] = relerr(ctx.download("RHS_O"), rhs_o_ref)
    if cfg["bctype"][0] == 11 and cfg["bctype"][1] == 1:      # body-fitted cylinder: wall areas and pressure / viscous forces (momentum.c:822-849)
        ref.set_owned("P", fields["p"])
        ref.global_to_local("P", "lP")
        ref.view("RHS_o")[...] = 0
        ref.Formfunction_2("RHS_o", 1.0)
        ctx.upload("P", fields["p"])
        err["cylinder_forces"] = relerr(ctx.cylinder_forces(), ref.cylinder_forces())
    if cfg["flags"].get("viscosity_wallmodel"):      # Cabot wall law: friction velocity at the first cells (momentum.c:1150)
        err["Ustar"] = relerr(ctx.download("USTAR")[:, 1], np.array(ref.owned("lUstar"))[:, 1])
    # one Krylov-iteration residual
    x = krylov_x(fields["ucont"])
    ref.new_vec("X", 3, False)
    ref.new_vec("F", 3, False)
    ref.view("X")[...] = x
    ref.FormFunction_SNES("X", "F")
    f_ref = np.array(ref.view("F"))
    f_dev = ctx.FormFunction_SNES(x)
    err["FormFunction_SNES"] = relerr(f_dev, f_ref)
    err["FormFunction_SNES_zero_pattern"] = float(np.count_nonzero((f_dev == 0) != (f_ref == 0)))
    err["SNES_ucat"] = relerr(ctx.download("UCAT"), ref.owned("Ucat"))
    # Pressure_Gradient (momentum.c:203-439) — last, because it overwrites dP.  k-periodic runs add the
    # mean-flux forcing (mean_k_flux - inlet_flux) / dt / mean_k_area per unit dz (:399-411)
    if "p" in fields:
        mkf, mka, infl = 0.3, 2.0, 0.1
        refdrv.set_global("inlet_flux", infl)
        ref.set_owned("P", fields["p"])
        ref.new_vec("dPg", 3, False)
        ref.Pressure_Gradient("dPg", mkf, mka)
        ctx.upload("P", fields["p"])
        ctx.Pressure_Gradient((mkf - infl) / cfg["dt"] / mka if (cfg["flags"].get("kk_periodic") or cfg["flags"].get("k_periodic")) else 0.0)
        err["Pressure_Gradient"] = relerr(ctx.download("DP"), ref.view("dPg"))
        err["Pressure_Gradient_P"] = relerr(ctx.download("P"), ref.owned("P"))
    ucat_before_projection = np.array(ref.owned("Ucat"))
    # UpdatePressure + Projection (poisson.c:3137, 2700), called back to back after the Poisson solve
    # (solvers.c:662-663): Phi stands in for the solver's pressure correction
    if "p" in fields:
        rng = np.random.default_rng(23)
        phi = 0.05 * rng.uniform(-1, 1, fields["p"].shape)
        phi = 0.25 * np.roll(phi, 1, 0) + 0.5 * phi + 0.25 * np.roll(phi, -1, 0)
        st = 0.9
        ref.set_owned("P", fields["p"]); ref.global_to_local("P", "lP")
        ref.set_owned("Phi", phi); ref.global_to_local("Phi", "lPhi")
        ref.set_owned("Ucont", fields["ucont"]); ref.global_to_local("Ucont", "lUcont")
        ref.UpdatePressure()
        ref.Projection(st)
        ctx.upload("P", fields["p"]); ctx.upload("PHI", phi); ctx.upload("UCONT", fields["ucont"])
        ctx.upload("UCAT", np.array(ucat_before_projection))
        ctx.UpdatePressure()
        err["UpdatePressure_P"] = relerr(ctx.download("P"), ref.owned("P"))
        err["UpdatePressure_Phi"] = relerr(ctx.download("PHI"), ref.owned("Phi"))
        ctx.Projection(st, refdrv.get_global("poisson_threshold"))
        err["Projection_ucont"] = relerr(ctx.download("UCONT"), ref.owned("Ucont"))
        ctx.Contra2Cart()                           # the reference's Projection ends with it (poisson.c:3049)
        err["Projection_lucont"] = relerr(ctx.download("UCONT"), ref.owned("lUcont"))
        err["Projection_ucat"] = relerr(ctx.download("UCAT"), ref.owned("Ucat"))
    # actuator forcing (rotor_model.c:3668, 2937) — after everything else: it overwrites F_eul
    act = make_actuator(cfg, xyz)
    ucat_now = np.array(ref.owned("Ucat"))
    ctx.upload("UCAT", ucat_now)
    ref.global_to_local("Ucat", "lUcat")
    err["Calc_U_lagr"] = relerr(ctx.Calc_U_lagr([act])[0], ref.Calc_U_lagr(act))
    for df in (0, 7, 10):
        ref.view("lF_eul")[...] = 0
        ref.Calc_F_eul(act, df)
        ctx.Calc_F_eul([act], df=df)
        err["Calc_F_eul_df%d" % df] = relerr(ctx.download("F_EUL"), ref.view("F_eul"))
    if verbose:
        for k, v in err.items():
            print("%-32s %.3e" % (k, v))
    ctx.close()
    return err


FIELDS_IN = (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"),
             ("ucont_rm1", "UCONT_RM1"), ("rhs_o", "RHS_O"), ("dp", "DP"), ("f_eul", "F_EUL"))


def run_path(ctx, x):
    out = {}
    ctx.Contra2Cart()
    ctx.Compute_Smagorinsky_Constant_1()
    ctx.Compute_eddy_viscosity_LES()
    out["F"] = ctx.FormFunction_SNES(x)
    for n in ("UCAT", "CS", "NU_T", "UCONT", "CSI", "AJ"):
        out[n] = ctx.download(n)
    # the same unit through the single fused entry point (its own, shorter, ghost-refresh schedule)
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




def run_golden(cfg, gold, lib=None, device=0):
    """Same path as run_parity but checked against a committed fixture (tests/golden/*.npz)
    generated from the reference by tests/golden/make_golden.py."""
    pkg = load_package()
    xyz = gold["xyz"]
    met = dict(csi=gold["csi"], eta=gold["eta"], zet=gold["zet"], aj=gold["aj"])
    fields = pkg.cases.make_fields(cfg, met)
    ctx = dev_setup(cfg, xyz, fields, lib=lib, device=device)
    err = {}
    for nm, key in (("CSI", "csi"), ("ETA", "eta"), ("ZET", "zet"), ("AJ", "aj")):
        err["metrics_" + nm] = relerr(ctx.download(nm), met[key])
    ctx.Contra2Cart()
    err["Contra2Cart_ucat"] = relerr(ctx.download("UCAT"), gold["ucat"])
    if cfg["flags"].get("les"):
        ctx.Compute_Smagorinsky_Constant_1()
        err["Cs"] = relerr(ctx.download("CS"), gold["cs"])
        ctx.Compute_eddy_viscosity_LES()
        err["nu_t"] = relerr(ctx.download("NU_T"), gold["nu_t"])
    if "conv" in gold.files:
        ctx.Convection(); ctx.Viscous()
        err["Convection"] = relerr(ctx.download("CONV") * conv_defined(cfg, fields), gold["conv"])
        err["Viscous"] = relerr(ctx.download("VISC"), gold["visc"])
    ctx.IB_BC()
    err["IB_BC_ucont"] = relerr(ctx.download("UCONT"), gold["ucont_after_ibbc"])
    ctx.upload("RHS_O", np.zeros_like(fields["rhs_o"]))
    ctx.Formfunction_2("RHS_O", 1.0)
    err["Formfunction_2"] = relerr(ctx.download("RHS_O"), gold["formfunction2"])
    ctx.upload("RHS_O", fields["rhs_o"])
    f_dev = ctx.FormFunction_SNES(krylov_x(fields["ucont"]))
    err["FormFunction_SNES"] = relerr(f_dev, gold["snes_f"])
    err["SNES_ucat"] = relerr(ctx.download("UCAT"), gold["snes_ucat"])
    ctx.close()
    return err
