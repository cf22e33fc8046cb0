"""Drop-in boundary test: the product's host glue (vfs-wind_b200/host/vfs_petsc_glue.cpp)
exports the reference's own function names and signatures.  The SAME driver harness that runs the
reference objects (oracle/harness.cpp) is linked against the glue instead, and the UserCtx Vecs it
leaves behind are compared with the reference run.  CPU variant: glue -> test-only kernel
emulation; the -m gpu twin (test_gpu_glue_dropin.py) links the CUDA library."""
import ctypes as C
import os
import numpy as np
import pytest
import parity_common as pc

TOL = 1e-12


def run_dropin(refdrv, pkg, so_name, cfg):
    """Run the reference-named entry points through `so_name` (a glue build) and through the real
    reference, on identical UserCtx contents.  Returns relative errors."""
    so = os.path.join(pc.ROOT, "oracle", "_ref", so_name)
    if not os.path.exists(so) and so_name.endswith("_emu.so") and os.path.isdir("/root/reference/Source"):
        # the glue is linked against whichever kernel libraries exist when oracle/build_ref.py runs:
        # build the test-only emulation first, then relink
        import subprocess, sys
        import emu_loader
        emu_loader.build()
        subprocess.check_call([sys.executable, os.path.join(pc.ROOT, "oracle", "build_ref.py")], stdout=subprocess.DEVNULL)
    if not os.path.exists(so):
        pytest.skip(so_name + " not built")
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    # second driver instance bound to the glue library
    import importlib.util
    spec = importlib.util.spec_from_file_location("refdrv_glue", os.path.join(pc.ROOT, "oracle", "refdrv.py"))
    gd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gd)
    gd.SO = so
    gd._GLOBALS_JSON = os.path.join(pc.ROOT, "oracle", "_ref", "globals_glue_%s.json" % so_name.split("_")[1].split(".")[0])
    glue, _, _, _ = pc.ref_setup(cfg, gd)
    err = {}
    for nm in ("lCsi", "lEta", "lZet", "lAj", "lICsi", "lJEta", "lKZet", "lIAj", "lKAj"):
        err["FormMetrics_" + nm] = pc.relerr(glue.owned(nm)[1:-1, 1:-1, 1:-1], ref.owned(nm)[1:-1, 1:-1, 1:-1])
    # side outputs of FormMetrics read by host code outside the path (metrics.c:89-107, 420-497)
    for nm in ("Cent", "lCent", "GridSpace", "lGridSpace"):
        err["FormMetrics_" + nm] = pc.relerr(glue.view(nm), ref.view(nm))
    ref.Contra2Cart(); glue.Contra2Cart()
    err["Contra2Cart_Ucat"] = pc.relerr(glue.owned("Ucat"), ref.owned("Ucat"))
    err["Contra2Cart_lUcat_ghosts"] = pc.relerr(glue.view("lUcat"), ref.view("lUcat"))
    # The host rewrites Ucat at IB / solid cells and boundary nodes between two calls (FormBCS, ibm_interpolation_advanced,
    # implicitsolver.c:4406-4444) WITHOUT the once-per-step invalidate: the reference keeps those values where no rule
    # touches them, so must the glue.
    rng = np.random.default_rng(7)
    nvm = np.array(ref.owned("Nvert")) > 0.1
    bnd = np.zeros(nvm.shape, bool)
    bnd[0] = bnd[-1] = True; bnd[:, 0] = bnd[:, -1] = True; bnd[:, :, 0] = bnd[:, :, -1] = True
    sel = nvm | bnd
    bump = 0.01 * rng.uniform(-1, 1, np.array(ref.owned("Ucat")).shape)
    for d in (ref, glue):
        u = np.array(d.owned("Ucat")); u[sel] += bump[sel]
        d.set_owned("Ucat", u)
        d.Contra2Cart()
    err["Contra2Cart_2nd_call_Ucat"] = pc.relerr(glue.owned("Ucat"), ref.owned("Ucat"))
    err["Contra2Cart_2nd_call_lUcat"] = pc.relerr(glue.view("lUcat"), ref.view("lUcat"))
    if any(b in (-1, -2) for b in cfg["bctype"][:4]):
        err["Contra2Cart_lUstar"] = pc.relerr(np.array(glue.owned("lUstar"))[1:-1, 1:-1, 1:-1], np.array(ref.owned("lUstar"))[1:-1, 1:-1, 1:-1])
    ref.Compute_Smagorinsky_Constant_1(); glue.Compute_Smagorinsky_Constant_1()
    err["lCs"] = pc.relerr(glue.owned("lCs"), ref.owned("lCs"))
    ref.Compute_eddy_viscosity_LES(); glue.Compute_eddy_viscosity_LES()
    err["lNu_t"] = pc.relerr(glue.owned("lNu_t"), ref.owned("lNu_t"))
    for d in (ref, glue):      # legacy explicit-solver terms, Vec arguments as in timeadvancing1.c:75-76
        d.new_vec("Conv", 3, False); d.new_vec("Visc", 3, False)
        d.Convection("Conv"); d.Viscous("Visc")
    ok = pc.conv_defined(cfg, fields)
    err["Convection"] = pc.relerr(np.array(glue.view("Conv")) * ok, np.array(ref.view("Conv")) * ok)
    err["Viscous"] = pc.relerr(glue.view("Visc"), ref.view("Visc"))
    ref.IB_BC(); glue.IB_BC()
    # Nodes lying on two or more domain-boundary planes are excluded: there IB_BC's component-wise
    # periodic copy (momentum.c:2210-2221) reads ghost images that the reference has not refreshed
    # since Contra2Cart rewrote their sources, while the glue re-uploads lUcont (fresh ghosts).
    # Those edge values feed nothing (Rhs is zeroed on boundary planes); the on-device sequence
    # used by FormFunction_SNES reproduces them exactly (tests/parity_common.py IB_BC_ucont).
    a, b = np.array(glue.owned("lUcont")), np.array(ref.owned("lUcont"))
    mz_, my_, mx_ = a.shape[:3]
    kk, jj, ii = np.meshgrid(np.arange(mz_), np.arange(my_), np.arange(mx_), indexing="ij")
    nb = ((kk == 0) | (kk == mz_ - 1)).astype(int) + ((jj == 0) | (jj == my_ - 1)) + ((ii == 0) | (ii == mx_ - 1))
    err["IB_BC_lUcont"] = pc.relerr(a[nb < 2], b[nb < 2])
    err["IB_BC_Nvert_mismatches"] = float(np.count_nonzero(np.array(glue.owned("Nvert")) != np.array(ref.owned("Nvert")))
                                          + np.count_nonzero(np.array(glue.owned("lNvert")) != np.array(ref.owned("lNvert"))))
    act = pc.make_actuator(cfg, xyz)                 # Calc_U_lagr / Calc_F_eul (rotor_model.c:2937, 3668) with the IBMNodes arrays
    err["Calc_U_lagr"] = pc.relerr(glue.Calc_U_lagr(act), ref.Calc_U_lagr(act))
    for d in (ref, glue):
        d.view("lF_eul")[...] = 0
        d.Calc_F_eul(act, 10)
    err["Calc_F_eul"] = pc.relerr(glue.view("F_eul"), ref.view("F_eul"))
    err["Calc_F_eul_lF_eul"] = pc.relerr(glue.view("lF_eul"), ref.view("lF_eul"))
    for d in (ref, glue):
        d.set_owned("F_eul", fields["f_eul"]); d.global_to_local("F_eul", "lF_eul")
    for d, drv in ((ref, refdrv), (glue, gd)):      # Pressure_Gradient (momentum.c:203), k-periodic mean-flux forcing included
        drv.set_global("inlet_flux", 0.1)
        d.set_owned("P", fields["p"])
        d.new_vec("dPg", 3, False)
        d.Pressure_Gradient("dPg", 0.3, 2.0)
    err["Pressure_Gradient"] = pc.relerr(glue.view("dPg"), ref.view("dPg"))
    err["Pressure_Gradient_lP"] = pc.relerr(glue.view("lP"), ref.view("lP"))
    for d in (ref, glue):
        d.view("RHS_o")[...] = 0
        d.Formfunction_2("RHS_o", 1.0)
    err["Formfunction_2_RHS_o"] = pc.relerr(glue.owned("RHS_o"), ref.owned("RHS_o"))
    if cfg["bctype"][0] == 11 and cfg["bctype"][1] == 1:      # the wall-force sums Formfunction_2 leaves in the UserCtx (momentum.c:822-849), read by main.c:1269
        err["cylinder_forces"] = pc.relerr(glue.cylinder_forces(), ref.cylinder_forces())
    x = fields["ucont"] * (1.0 + 1e-3 * np.cos(np.arange(fields["ucont"].size).reshape(fields["ucont"].shape)))
    for d in (ref, glue):
        d.new_vec("X", 3, False); d.new_vec("F", 3, False)
        d.view("X")[...] = x
        gd.lib().vfs_glue_invalidate(C.c_void_p(glue.u)) if d is glue else None
        d.FormFunction_SNES("X", "F")
    err["FormFunction_SNES_F"] = pc.relerr(glue.view("F"), ref.view("F"))
    # side effects of FormFunction_SNES on the UserCtx Vecs (momentum.c:2240-2295), read by the reference right after
    # SNESSolve (implicitsolver.c:4360-4376, 4440): materialised by the documented vfs_glue_sync_state call
    gd.lib().vfs_glue_sync_state(C.c_void_p(glue.u))
    for nm in ("Ucont", "Ucat", "Nvert"):
        err["SNES_side_effect_" + nm] = pc.relerr(glue.view(nm), ref.view(nm))
    err["SNES_side_effect_lUcat"] = pc.relerr(glue.view("lUcat"), ref.view("lUcat"))
    err["SNES_side_effect_lNvert"] = pc.relerr(glue.owned("lNvert"), ref.owned("lNvert"))
    a, b = np.array(glue.owned("lUcont")), np.array(ref.owned("lUcont"))
    err["SNES_side_effect_lUcont"] = pc.relerr(a[nb < 2], b[nb < 2])
    # the same with the eager switch: every evaluation mirrors its side effects, no explicit sync
    gd.lib().vfs_glue_set_eager(1)
    x2 = x * (1.0 + 1e-3)
    for d in (ref, glue):
        d.view("X")[...] = x2
        d.FormFunction_SNES("X", "F")
    gd.lib().vfs_glue_set_eager(0)
    err["eager_F"] = pc.relerr(glue.view("F"), ref.view("F"))
    for nm in ("Ucont", "Ucat"):
        err["eager_side_effect_" + nm] = pc.relerr(glue.view(nm), ref.view(nm))
    # after the Poisson solve: UpdatePressure + Projection (poisson.c:3137, 2700; solvers.c:662-663), every Vec they leave behind
    phi = 0.05 * np.random.default_rng(23).uniform(-1, 1, fields["p"].shape)
    for d in (ref, glue):
        d.set_owned("P", fields["p"]); d.global_to_local("P", "lP")
        d.set_owned("Phi", phi); d.global_to_local("Phi", "lPhi")
        d.set_owned("Ucont", fields["ucont"]); d.global_to_local("Ucont", "lUcont")
        d.UpdatePressure()
        d.Projection(0.9)
    for nm in ("P", "lP", "Phi", "lPhi", "Ucont", "Ucat", "lUcat"):
        err["Projection_" + nm] = pc.relerr(glue.view(nm), ref.view(nm))
    # lUcont: owned part only — Contra2Cart_2 rewrites lUcont's periodic boundary nodes in place (rhs.c:129-156) and the
    # reference leaves the ghost images of those nodes stale, the glue refreshes them
    err["Projection_lUcont"] = pc.relerr(glue.owned("lUcont"), ref.owned("lUcont"))
    gd.lib().vfs_glue_release(C.c_void_p(glue.u))
    return err


def variant_cfg(pkg, dims):
    """Body-fitted cylinder boundary (bctype 11) with the skew-symmetric form and the Clark mixed model switched on."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], *dims)
    cfg["flags"] = dict(cfg["flags"], ii_periodic=0, kk_periodic=0, skew=1, clark=1)
    cfg["bctype"] = [11, 1, 1, 1, 5, 4]
    cfg["z_shift"] = 1.4
    return cfg


@pytest.mark.parametrize("name,dims", [("c2_box256", (13, 11, 15)), ("c3_turbine", (19, 15, 17)), ("wallfn", (13, 11, 15)), ("variants", (17, 13, 15)), ("legacy_periodic", (13, 11, 15))])
def test_glue_dropin_emulated(pkg, refdrv, name, dims):
    import emu_loader
    emu_loader.build()
    if name == "variants":
        cfg = variant_cfg(pkg, dims)
    elif name == "legacy_periodic":      # i/k periodic through the legacy switches: the host DA is NOT periodic, its local Vecs have no wrap ghosts
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c2_box256"], *dims)
        cfg["flags"] = dict(cfg["flags"], ii_periodic=0, kk_periodic=0, i_periodic=1, k_periodic=1)
    elif name == "wallfn":       # wall-function sides, first time step: IB_BC rewrites lNvert / Nvert on the host too
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c2_box256"], *dims)
        cfg["flags"] = dict(cfg["flags"], ti=5, tistart=5, roughness_size=2.e-4)
        cfg["bctype"] = [100, 100, -1, -2, 100, 100]
    else:
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = run_dropin(refdrv, pkg, "libvfsglue_emu.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def run_periodic_turbines(refdrv, pkg, so_name):
    """Calc_U_lagr on a 2 x 1 x 2 periodic turbine array in a moving frame (rotor_model.c:3062-3138): the first / last
    objects of the periodic directions share their sums, the frame velocity is added — host bookkeeping the glue does
    after the device interpolation."""
    import importlib.util
    so = os.path.join(pc.ROOT, "oracle", "_ref", so_name)
    if not os.path.exists(so):
        pytest.skip(so_name + " not built")
    spec = importlib.util.spec_from_file_location("refdrv_glue3", os.path.join(pc.ROOT, "oracle", "refdrv.py"))
    gd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gd)
    gd.SO = so
    gd._GLOBALS_JSON = os.path.join(pc.ROOT, "oracle", "_ref", "globals_glue_%s.json" % so_name.split("_")[1].split(".")[0])
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 25, 17, 21)
    cfg["flags"] = dict(cfg["flags"], ii_periodicWT=1, kk_periodicWT=1, Nx_WT=2, Ny_WT=1, Nz_WT=2, Sx_WT=0.8, Sy_WT=1.0, Sz_WT=1.1,
                        MoveFrame=1, u_frame=0.05, v_frame=-0.02, w_frame=0.3)
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    glue, _, _, _ = pc.ref_setup(cfg, gd)
    acts = [pc.make_actuator(cfg, xyz, seed=11 + q, n=12) for q in range(4)]
    centres = np.array([[0.3, 0.4, 0.5], [1.1, 0.4, 0.5], [0.3, 0.4, 1.6], [1.1, 0.4, 1.6]])      # object -> array slot (i, k): (0,0) (1,0) (0,1) (1,1)
    for d in (ref, glue):
        d.Contra2Cart()
    a, b = ref.Calc_U_lagr_multi(acts, centres), glue.Calc_U_lagr_multi(acts, centres)
    err = {"U_lagr_object_%d" % q: pc.relerr(b[q], a[q]) for q in range(4)}
    assert np.array_equal(a[1], a[0]) and np.array_equal(a[2], a[0])      # images of the first object, as the reference leaves them
    gd.lib().vfs_glue_release(C.c_void_p(glue.u))
    return err


def test_glue_periodic_turbine_array_emulated(pkg, refdrv):
    import emu_loader
    emu_loader.build()
    err = run_periodic_turbines(refdrv, pkg, "libvfsglue_emu.so")
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def run_two_time_steps(refdrv, pkg, so_name, cfg):
    """Flow_Solver's call sequence (solvers.c:365-688) for two time steps, on the reference objects and on the glue, with
    the documented once-per-step `vfs_glue_invalidate`: LES block, Pressure_Gradient, RHS_o = Formfunction_2(U), three
    residual evaluations standing in for the Krylov iterations, a stand-in pressure correction, UpdatePressure +
    Projection, then the host-side end-of-step bookkeeping (Ucont_o <- Ucont, lUcat_old <- Ucat).  Catches stale cached
    constants: every Vec either side leaves behind is compared after each step."""
    import importlib.util
    so = os.path.join(pc.ROOT, "oracle", "_ref", so_name)
    if not os.path.exists(so):
        pytest.skip(so_name + " not built")
    spec = importlib.util.spec_from_file_location("refdrv_glue4", os.path.join(pc.ROOT, "oracle", "refdrv.py"))
    gd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gd)
    gd.SO = so
    gd._GLOBALS_JSON = os.path.join(pc.ROOT, "oracle", "_ref", "globals_glue_%s.json" % so_name.split("_")[1].split(".")[0])
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    glue, _, _, _ = pc.ref_setup(cfg, gd)
    rng = np.random.default_rng(31)
    err = {}
    act = pc.make_actuator(cfg, xyz)
    for d in (ref, glue):
        d.set_owned("P", fields["p"]); d.global_to_local("P", "lP")
        d.new_vec("X", 3, False); d.new_vec("F", 3, False)
    for step in range(2):
        phi = 0.02 * rng.uniform(-1, 1, fields["p"].shape)
        pert = [1e-3 * rng.uniform(-1, 1, fields["ucont"].shape) for _ in range(3)]
        for d, drv in ((ref, refdrv), (glue, gd)):
            drv.set_global("ti", 10 + step)
            if d is glue:
                gd.lib().vfs_glue_invalidate(C.c_void_p(glue.u))
            d.global_to_local("Ucont", "lUcont")
            d.Contra2Cart(); d.Compute_Smagorinsky_Constant_1(); d.Compute_eddy_viscosity_LES()      # solvers.c:365-370
            d.Pressure_Gradient("dP", 0.3, 2.0)                                                        # :438 (into user->dP, as the reference calls it)
            if cfg["flags"].get("rotor_model"):                                                        # :491-534: velocity at the blades -> (host aerodynamics) -> forcing
                ul = d.Calc_U_lagr(act)
                act_step = dict(act, F_lagr=act["F_lagr"] * (1.0 + 0.1 * step) - 0.05 * ul)
                d.view("lF_eul")[...] = 0
                d.Calc_F_eul(act_step, 10)
            d.view("RHS_o")[...] = 0
            d.Formfunction_2("RHS_o", 1.0)                                                             # :629
            # (no second invalidate: dP and RHS_o were produced by the glue, the device copies are the fresh ones)
            u = np.array(d.owned("Ucont"))
            for q in range(3):                                                                         # :631 (SNESSolve)
                d.view("X")[...] = u * (1.0 + pert[q])
                d.FormFunction_SNES("X", "F")
                u = u + 0.1 * cfg["dt"] * np.array(d.view("F"))
            if d is glue:
                gd.lib().vfs_glue_sync_state(C.c_void_p(glue.u))
            d.set_owned("Ucont", u); d.global_to_local("Ucont", "lUcont")
            d.set_owned("Phi", phi); d.global_to_local("Phi", "lPhi")                                  # :652 (Poisson solve, host)
            d.UpdatePressure(); d.Projection(1.0)                                                      # :662-663
        for nm in ("Ucont", "Ucat", "lUcat", "P", "lP", "RHS_o", "dP", "F_eul"):
            err["step%d_%s" % (step, nm)] = pc.relerr(glue.view(nm), ref.view(nm))
        for nm in ("lCs", "lNu_t", "lUcont"):
            err["step%d_%s" % (step, nm)] = pc.relerr(glue.owned(nm), ref.owned(nm))
        err["step%d_F" % step] = pc.relerr(glue.view("F"), ref.view("F"))
        for d in (ref, glue):                         # end of the step: Ucont_o <- Ucont, lUcat_old <- Ucat (solvers.c / main.c)
            d.set_owned("Ucont_o", np.array(d.owned("Ucont")))
            d.set_owned("lUcat_old", np.array(d.owned("Ucat"))); d.wrap_fill("lUcat_old")
    gd.lib().vfs_glue_release(C.c_void_p(glue.u))
    return err


@pytest.mark.parametrize("name,dims", [("c2_box256", (13, 11, 15)), ("c3_turbine", (19, 15, 17)), ("variants", (17, 13, 15)), ("legacy_periodic", (13, 11, 15))])
def test_glue_two_time_steps_emulated(pkg, refdrv, name, dims):
    import emu_loader
    emu_loader.build()
    if name == "variants":
        cfg = variant_cfg(pkg, dims)
    elif name == "legacy_periodic":
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c2_box256"], *dims)
        cfg["flags"] = dict(cfg["flags"], ii_periodic=0, kk_periodic=0, i_periodic=1, k_periodic=1)
    else:
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = run_two_time_steps(refdrv, pkg, "libvfsglue_emu.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= 1e-11)}
    assert not bad, bad


def run_glue_solver(refdrv, pkg, so_name, cfg):
    """vfs_glue_snes_solve — the one-line replacement of SNESSolve in Implicit_MatrixFree (implicitsolver.c:4299) —
    against the numpy restatement of the PETSc algorithms driving the REFERENCE residual; also the UserCtx Vecs it
    leaves behind (Ucont, Ucat) against the reference's after its last residual evaluation."""
    import importlib.util
    import newton_krylov_ref as nk
    so = os.path.join(pc.ROOT, "oracle", "_ref", so_name)
    if not os.path.exists(so):
        pytest.skip(so_name + " not built")
    spec = importlib.util.spec_from_file_location("refdrv_glue2", os.path.join(pc.ROOT, "oracle", "refdrv.py"))
    gd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gd)
    gd.SO = so
    gd._GLOBALS_JSON = os.path.join(pc.ROOT, "oracle", "_ref", "globals_glue_%s.json" % so_name.split("_")[1].split(".")[0])
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    glue, _, _, _ = pc.ref_setup(cfg, gd)
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    for nm in ("RHS_o", "dP", "F_eul"):
        ref.view(nm)[...] = 0
    ref.Contra2Cart(); ref.Compute_Smagorinsky_Constant_1(); ref.Compute_eddy_viscosity_LES()
    ref.view("X")[...] = fields["ucont"]
    ref.FormFunction_SNES("X", "F")
    live = (np.array(ref.view("F")) != 0).astype(float)          # see tests/solver_common.py
    for key, nm in (("rhs_o", "RHS_o"), ("dp", "dP"), ("f_eul", "F_eul")):
        for d in (ref, glue):
            d.set_owned(nm, fields[key] * live)
    for d in (ref, glue):
        d.set_owned("Ucont", fields["ucont"]); d.global_to_local("Ucont", "lUcont")
        d.set_owned("Ucat", fields["ucat"]); d.global_to_local("Ucat", "lUcat")
        d.Contra2Cart(); d.Compute_Smagorinsky_Constant_1(); d.Compute_eddy_viscosity_LES()

    def residual(x):
        ref.view("X")[...] = x
        ref.FormFunction_SNES("X", "F")
        return np.array(ref.view("F"))
    u_ref, info_ref = nk.snes_tr(residual, fields["ucont"])
    residual(u_ref)                                               # the reference's Vecs after its last evaluation
    L = gd.lib()
    sp = pkg.capi.VfsSolverParams(); info = pkg.capi.VfsSolverInfo()
    emu_or_cuda = C.CDLL(os.path.join(pc.ROOT, "tests", "emu", "libvfs_emu.so") if "emu" in so_name else pkg.capi.LIB_PATH)
    emu_or_cuda.vfs_solver_defaults(C.byref(sp))
    glue.new_vec("U", 3, False)
    glue.view("U")[...] = fields["ucont"]
    L.vfs_glue_invalidate(C.c_void_p(glue.u))
    L.vfs_glue_snes_solve.restype = C.c_double
    fn = L.vfs_glue_snes_solve(C.c_void_p(glue.u), C.c_void_p(glue.vec("U")), C.byref(sp), C.byref(info))
    err = {"U": pc.relerr(glue.view("U"), u_ref), "fnorm": abs(fn - info_ref["fnorm_history"][-1]) / info_ref["fnorm_history"][0],
           "krylov_its_differ": float([info.ksp_its_history[q] for q in range(info.n_history - 1)] != info_ref["ksp_its_history"]),
           "Ucont": pc.relerr(glue.view("Ucont"), ref.view("Ucont")), "Ucat": pc.relerr(glue.view("Ucat"), ref.view("Ucat"))}
    L.vfs_glue_release(C.c_void_p(glue.u))
    return err


def test_glue_snes_solve_emulated(pkg, refdrv):
    import emu_loader
    emu_loader.build()
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15)
    err = run_glue_solver(refdrv, pkg, "libvfsglue_emu.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= 1e-10)}
    assert not bad, bad
