"""CPU (no GPU) checks of the kernel LOGIC: the kernel functors compiled for the host
(tests/emu, -DVFS_EMU, test-only) against the oracle.  The -m gpu tests repeat these through the
real CUDA library."""
import numpy as np
import pytest
import parity_common as pc
import emu_loader

TOL = 1e-12


@pytest.fixture(scope="module")
def emu(pkg):
    return emu_loader.load(pkg.capi)


@pytest.mark.parametrize("name,dims", [("c2_box256", (13, 11, 15)), ("c3_turbine", (21, 17, 19)), ("c1_test10", (14, 10, 12))])
def test_emulated_path_matches_reference(pkg, refdrv, emu, name, dims):
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, lib=emu)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_emulated_flag_variants(pkg, refdrv, emu):
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15)
    for extra in (dict(second_order=1), dict(laplacian=1), dict(immersed=3), dict(les=1), dict(les=0), dict(testfilter_ik=1),
                  dict(kk_periodic=1, ii_periodic=0), dict(jj_periodic=1)):
        cfg = dict(base)
        cfg["flags"] = dict(base["flags"], **extra)
        if extra.get("kk_periodic") or extra.get("jj_periodic"):
            cfg["bctype"] = [1, 1, 1, 1, 100, 100] if extra.get("kk_periodic") else [100, 100, 100, 100, 5, 4]
        err = pc.run_parity(cfg, refdrv, lib=emu)
        assert err.pop("FormFunction_SNES_zero_pattern") == 0, extra
        bad = {k: v for k, v in err.items() if not (v <= TOL)}
        assert not bad, (extra, bad)


@pytest.mark.parametrize("name,dims", [("c2_box256", (45, 30, 27)), ("c3_turbine", (41, 33, 25))])
def test_emulated_fused_residual_option(pkg, refdrv, emu, name, dims):
    """Option 0 = 2: the fully fused residual marching kernel (interior) + staged boundary slabs, several
    tiles and k-chunks wide, against the oracle."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, lib=emu, options={0: 2})
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_emulated_staged_option(pkg, refdrv, emu):
    """Option 0 = 0: the one-thread-per-cell staged kernels only (the literal restatement)."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 21, 17, 19)
    err = pc.run_parity(cfg, refdrv, lib=emu, options={0: 0})
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("bctype,extra", [
    ([100, 100, -1, -2, 100, 100], dict(ii_periodic=1, kk_periodic=1, immersed=0, ti=5, tistart=5, roughness_size=2.e-4)),   # first step: IB_BC marks the first cells
    ([-1, -2, -2, -1, 5, 4], dict(ii_periodic=0, kk_periodic=0, immersed=0, roughness_size=1.e-3)),                           # all four i/j sides
    ([100, 100, -1, 10, 5, 4], dict(ii_periodic=1, kk_periodic=0)),                                                           # with immersed bodies (c3)
])
def test_emulated_wall_function_boundaries(pkg, refdrv, emu, bctype, extra):
    """bctype -1 (Cabot) / -2 (rough log law): first-cell velocities of Contra2Cart_2 (rhs.c:311-440), u_tau, IB_BC's
    first-step nvert marking and wall-face flux zeroing (momentum.c:2048-2074, 2169-2189)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], **extra)
    cfg["bctype"] = bctype
    err = pc.run_parity(cfg, refdrv, lib=emu)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    assert err.pop("IB_BC_nvert_mismatches") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


# the branches VERDICT round 1 listed as rejected: WENO3 convection (inviscid / levelset_weno = 5, momentum.c:754-770), the
# skew-symmetric form (:789-800, 1638-1651), the Clark mixed model in the viscous flux (:904-923) and in the Germano
# identity (les.c:420-428, 497-556, 656), each alone and combined with periodic seams / second order / the box filter
VARIANT_FLAGS = [dict(inviscid=1), dict(levelset_weno=5), dict(skew=1), dict(skew=1, second_order=1), dict(clark=1), dict(clark=1, les=0),
                 dict(clark=1, testfilter_ik=1), dict(skew=1, clark=1, kk_periodic=1, ii_periodic=0), dict(skew=1, jj_periodic=1, levelset_weno=5),
                 dict(inviscid=1, immersed=3), dict(wallfunction=2)]      # (wallfunction = 2: nu_t zeroed next to IB nodes, les.c:1211)


def _variant_cfg(base, extra):
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], **extra)
    if extra.get("kk_periodic") or extra.get("jj_periodic"):
        cfg["bctype"] = [1, 1, 1, 1, 100, 100] if extra.get("kk_periodic") else [100, 100, 100, 100, 5, 4]
    return cfg


@pytest.mark.parametrize("extra", VARIANT_FLAGS)
def test_emulated_weno_skew_clark_variants(pkg, refdrv, emu, extra):
    cfg = _variant_cfg(pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15), extra)
    err = pc.run_parity(cfg, refdrv, lib=emu)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_emulated_cylinder_inflow_boundary(pkg, refdrv, emu):
    """bctype[0] = 11 (body-fitted cylinder, rhs.c:626-634): the inflow mirror at the i = 0 nodes whose cell centre has
    z <= 0 (the grid is shifted so that about half of them do), and the wall-force diagnostics Formfunction_2 accumulates
    over the i = mx-2 faces (momentum.c:570-579, 822-849)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], ii_periodic=0, kk_periodic=0)
    cfg["bctype"] = [11, 1, 1, 1, 5, 4]
    cfg["z_shift"] = 1.4
    err = pc.run_parity(cfg, refdrv, lib=emu)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    assert "cylinder_forces" in err
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("extra,bctype", [
    (dict(ii_periodic=0, kk_periodic=0, i_periodic=1, k_periodic=1), [100, 100, 1, 10, 100, 100]),
    (dict(ii_periodic=0, kk_periodic=0, i_periodic=1), [100, 100, 1, 10, 5, 4]),
    (dict(ii_periodic=0, kk_periodic=0, j_periodic=1, k_periodic=1, skew=1), [1, 1, 100, 100, 100, 100]),
])
def test_emulated_legacy_periodic_switches(pkg, refdrv, emu, extra, bctype):
    """The legacy i/j/k_periodic switches (explicit index remaps m-2 / 1 / m-3 / 2 on a non-periodic DA, e.g.
    momentum.c:644-651, 708-711, 1575-1601; single rank in the reference): the same ghost images as the DA wrap, except that
    IB_BC's component copies read the interior node live instead of its stale ghost image (momentum.c:2206-2211)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 17, 13, 15)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], **extra)
    cfg["bctype"] = bctype
    err = pc.run_parity(cfg, refdrv, lib=emu)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_emulated_fp_in_projection_bitwise(pkg, emu):
    """Option 12 (Fp evaluated inside the projection block program) against FpCell + Fp planes + Project: bitwise."""
    import numpy as np
    capi, cases = pkg.capi, pkg.cases
    for cfgname, dims in (("c2_box256", (37, 19, 23)), ("c3_turbine", (35, 21, 19))):
        cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        outs = []
        for val in (0, 1, 2):
            ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]), lib=emu)
            ctx.set_option(12, val)
            ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
            for k, n in pc.FIELDS_IN:
                ctx.upload(n, f[k])
            o = pc.run_path(ctx, pc.krylov_x(f["ucont"]))
            ctx.upload("RHS_O", f["rhs_o"]); ctx.Formfunction_2("RHS_O", 0.7)
            o["FF2"] = ctx.download("RHS_O")
            outs.append(o)
            ctx.close()
        for n in outs[0]:
            assert np.array_equal(outs[0][n], outs[1][n]), (cfgname, n)
            assert np.array_equal(outs[0][n], outs[2][n]), (cfgname, n, 'box')


@pytest.mark.parametrize("extra", [dict(i_homo_filter=1, k_homo_filter=1), dict(i_homo_filter=1), dict(j_homo_filter=1), dict(k_homo_filter=1)])
def test_emulated_homogeneous_cs_averaging(pkg, refdrv, emu, extra):
    """les.c:798-965: Cs from LM, MM averaged over the homogeneous direction(s) (channel-flow setting: i and k)."""
    for name, dims in (("c2_box256", (13, 11, 15)), ("c3_turbine", (17, 13, 15))):
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
        cfg["flags"] = dict(cfg["flags"], **extra)
        err = pc.run_parity(cfg, refdrv, lib=emu, legacy=False)
        assert err.pop("FormFunction_SNES_zero_pattern") == 0, extra
        bad = {k: v for k, v in err.items() if not (v <= TOL)}
        assert not bad, (extra, bad)


def test_emulated_restart_files_feed_the_path(pkg, emu, tmp_path):
    """SURVEY 8(f) row f4 end to end: a grid.dat (binary) and the restart file set of a time step are written in the
    reference's formats, read back with petsc_io and fed to a context — the unit's results are bitwise those of the
    context fed from memory."""
    capi, cases, io = pkg.capi, pkg.cases, pkg.petsc_io
    cfg = cases.scaled(cases.CONFIGS["c3_turbine"], 21, 15, 17)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    xyz = cases.make_grid(cfg)
    io.write_grid_dat(str(tmp_path / "grid.dat"), [xyz], binary=True)
    outs = []
    for from_disk in (False, True):
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]), lib=emu)
        ctx.upload("COOR", io.read_grid_dat(str(tmp_path / "grid.dat"), binary=True)[0] if from_disk else xyz)
        ctx.FormMetrics()
        if not from_disk:
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
            io.write_restart(str(tmp_path), 40, dict(UCONT=f["ucont"], UCAT=f["ucat"], P=f["p"], NVERT=f["nvert"]))
        for k, n in pc.FIELDS_IN:
            ctx.upload(n, f[k])
        if from_disk:      # as Ucont_Read does: Ucont_o <- Ucont, lUcat_old <- Ucat (main.c:420-428)
            for n, a in io.read_restart(str(tmp_path), 40, mx, my, mz).items():
                ctx.upload(n, a)
        else:
            ctx.upload("UCONT_O", f["ucont"]); ctx.upload("UCAT_OLD", f["ucat"]); ctx.upload("P", f["p"])
        ctx.rhs_les_fused()
        outs.append({n: ctx.download(n) for n in ("RHS", "UCAT", "CS", "NU_T")})
        ctx.close()
    for n in outs[0]:
        assert np.array_equal(outs[0][n], outs[1][n]), n


@pytest.mark.parametrize("name,dims", [("c2_box256", (17, 13, 15)), ("c3_turbine", (21, 17, 19))])
def test_emulated_state_carries_over_several_steps(pkg, refdrv, emu, name, dims):
    """Persistent device state across calls (Ucat is in/out, IB and boundary values carry over): four
    explicit pseudo-steps U <- U + 0.2 dt F(U), each = the LES block + one residual, on the reference and through the fused
    unit of the library, nothing re-uploaded but U.  The residual of every step must agree (the iterates therefore too)."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    ref, xyz, fields, met = pc.ref_setup(cfg, refdrv)
    ctx = pc.dev_setup(cfg, xyz, fields, lib=emu)
    ref.new_vec("X", 3, False); ref.new_vec("F", 3, False)
    x = np.array(fields["ucont"])
    for step in range(4):
        ref.set_owned("Ucont", x); ref.global_to_local("Ucont", "lUcont")
        ref.Contra2Cart(); ref.Compute_Smagorinsky_Constant_1(); ref.Compute_eddy_viscosity_LES()
        ref.view("X")[...] = x
        ref.FormFunction_SNES("X", "F")
        f_ref = np.array(ref.view("F"))
        ctx.upload("UCONT", x)
        ctx.rhs_les_fused()
        f_dev = ctx.download("RHS")
        assert pc.relerr(f_dev, f_ref) <= 1e-11, (step, pc.relerr(f_dev, f_ref))
        assert pc.relerr(ctx.download("UCAT"), ref.owned("Ucat")) <= 1e-12, step
        assert pc.relerr(ctx.download("NU_T"), ref.owned("lNu_t")) <= 1e-11, step
        x = x + 0.2 * cfg["dt"] * f_ref
    ctx.close()
