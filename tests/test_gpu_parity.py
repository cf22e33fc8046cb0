"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle on the same seeded
inputs.  Tolerance (BASELINE.json north_star): relative 1e-12 of the field's max norm for Rhs,
Ucat, Cs, nu_t; mask-driven zero patterns bit-exact."""
import numpy as np
import pytest
import parity_common as pc

TOL = 1e-12
pytestmark = pytest.mark.gpu

CASES = [
    ("c2_box256", (21, 17, 25)),      # ii+kk periodic, 4th-order central, dynamic Smagorinsky
    ("c2_box256", (40, 33, 37)),
    ("c3_turbine", (29, 21, 25)),     # IBM masks, QUICK at IB faces, inflow/outflow in k, F_eul
    ("c3_turbine", (45, 30, 41)),
    ("c1_test10", (24, 16, 20)),      # 2nd-order, laplacian, Cabot wall model at the j = 0 faces
]


@pytest.mark.parametrize("name,dims", CASES)
def test_path_matches_reference(pkg, refdrv, name, dims):
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_c1_on_its_shipped_grid(pkg, refdrv):
    """BASELINE.json configs[0]: Test_10_ChannelFlow_Retau3000 on its shipped 122 x 42 x 62-node grid (cases.make_grid
    reproduces the shipped xyz.dat to 1e-14, tests/test_cpu_formats.py), flags of its control.dat, wall model on:
    the whole path through the CUDA library against the oracle at the case's real size."""
    cfg = dict(pkg.cases.CONFIGS["c1_test10"])
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("name,dims", [("c2_box256", (127, 95, 79)), ("c3_turbine", (99, 69, 89))])
def test_multi_tile_sizes_match_reference(pkg, refdrv, name, dims):
    """Oracle parity of the marching / TMA kernels on grids several tiles wide and several k-chunks deep
    (128 x 96 x 80 nodes: 4 x 6 flux tiles, 5 x 10 LES-2 tiles), masks and F_eul included."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, device=0, legacy=False)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("name,dims", [("c2_box256", (70, 37, 45)), ("c3_turbine", (45, 30, 41))])
def test_fused_residual_option_matches_reference(pkg, refdrv, name, dims):
    """Option 0 = 2: the fully fused residual marching kernel (TMA ring, fluxes and Fp on chip) on the
    regular interior + staged boundary slabs."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, device=0, options={0: 2})
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_flag_variants(pkg, refdrv):
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 25, 19, 23)
    for extra in (dict(second_order=1), dict(laplacian=1), dict(immersed=3), dict(les=1), dict(les=0), dict(testfilter_ik=1),
                  dict(kk_periodic=1, ii_periodic=0), dict(jj_periodic=1)):
        cfg = dict(base)
        cfg["flags"] = dict(base["flags"], **extra)
        if extra.get("kk_periodic") or extra.get("jj_periodic"):
            cfg["bctype"] = [1, 1, 1, 1, 100, 100] if extra.get("kk_periodic") else [100, 100, 100, 100, 5, 4]
        err = pc.run_parity(cfg, refdrv, device=0)
        assert err.pop("FormFunction_SNES_zero_pattern") == 0, extra
        bad = {k: v for k, v in err.items() if not (v <= TOL)}
        assert not bad, (extra, bad)


@pytest.mark.parametrize("bctype,extra", [
    ([100, 100, -1, -2, 100, 100], dict(ii_periodic=1, kk_periodic=1, immersed=0, ti=5, tistart=5, roughness_size=2.e-4)),
    ([-1, -2, -2, -1, 5, 4], dict(ii_periodic=0, kk_periodic=0, immersed=0, roughness_size=1.e-3)),
    ([100, 100, -1, 10, 5, 4], dict(ii_periodic=1, kk_periodic=0)),
])
def test_wall_function_boundaries(pkg, refdrv, bctype, extra):
    """bctype -1 (Cabot) / -2 (rough log law) on i/j sides: first-cell velocities and u_tau (rhs.c:311-440), IB_BC's
    first-step nvert marking (bit-exact) and wall-face flux zeroing (momentum.c:2048-2074, 2169-2189)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 45, 30, 41)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], **extra)
    cfg["bctype"] = bctype
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    assert err.pop("IB_BC_nvert_mismatches") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("extra", [dict(i_homo_filter=1, k_homo_filter=1), dict(i_homo_filter=1), dict(j_homo_filter=1), dict(k_homo_filter=1)])
def test_homogeneous_cs_averaging(pkg, refdrv, extra):
    """les.c:798-965: Cs from LM, MM averaged over the homogeneous direction(s); device reductions in a fixed order."""
    for name, dims in (("c2_box256", (40, 33, 37)), ("c3_turbine", (45, 30, 41))):
        cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
        cfg["flags"] = dict(cfg["flags"], **extra)
        err = pc.run_parity(cfg, refdrv, device=0, legacy=False)
        assert err.pop("FormFunction_SNES_zero_pattern") == 0, extra
        bad = {k: v for k, v in err.items() if not (v <= TOL)}
        assert not bad, (extra, bad)


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
def test_weno_skew_clark_variants(pkg, refdrv, extra):
    cfg = _variant_cfg(pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 37, 25, 29), extra)
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_cylinder_inflow_boundary(pkg, refdrv):
    """bctype[0] = 11 (rhs.c:626-634) and the wall-force diagnostics of Formfunction_2 (momentum.c:822-849)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 37, 25, 29)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], ii_periodic=0, kk_periodic=0)
    cfg["bctype"] = [11, 1, 1, 1, 5, 4]
    cfg["z_shift"] = 1.4
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    assert "cylinder_forces" in err
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("extra,bctype", [
    (dict(ii_periodic=0, kk_periodic=0, i_periodic=1, k_periodic=1), [100, 100, 1, 10, 100, 100]),
    (dict(ii_periodic=0, kk_periodic=0, i_periodic=1), [100, 100, 1, 10, 5, 4]),
    (dict(ii_periodic=0, kk_periodic=0, j_periodic=1, k_periodic=1, skew=1), [1, 1, 100, 100, 100, 100]),
])
def test_legacy_periodic_switches(pkg, refdrv, extra, bctype):
    """The legacy i/j/k_periodic switches (explicit index remaps m-2 / 1 / m-3 / 2 on a non-periodic DA, e.g.
    momentum.c:644-651, 708-711, 1575-1601; single rank in the reference): the same ghost images as the DA wrap, except that
    IB_BC's component copies read the interior node live instead of its stale ghost image (momentum.c:2206-2211)."""
    base = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 37, 25, 29)
    cfg = dict(base)
    cfg["flags"] = dict(base["flags"], **extra)
    cfg["bctype"] = bctype
    err = pc.run_parity(cfg, refdrv, device=0)
    assert err.pop("FormFunction_SNES_zero_pattern") == 0
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_unsupported_flags_fail_loudly(pkg):
    capi = pkg.capi
    p = capi.make_params(16, 16, 16, dict(les=2, levelset=1), 100.0, 1e-3, [1] * 6)
    with pytest.raises(capi.VfsError):
        capi.VfsContext(p)


def test_roundtrip_and_idempotence(pkg):
    """Size-independent properties at a larger size: upload/download round trip is exact and a
    second residual evaluation on the same X is bitwise identical (no hidden state drift)."""
    capi, cases = pkg.capi, pkg.cases
    cfg = cases.scaled(cases.CONFIGS["c2_box256"], 95, 63, 79)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
    xyz = cases.make_grid(cfg)
    ctx.upload("COOR", xyz)
    assert np.array_equal(ctx.download("COOR"), xyz)
    ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    assert np.all(met["aj"][1:-1, 1:-1, 1:-1] > 0)
    f = cases.make_fields(cfg, met)
    for k, n in (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"), ("rhs_o", "RHS_O"), ("dp", "DP")):
        ctx.upload(n, f[k])
    ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
    F1 = ctx.FormFunction_SNES(f["ucont"])
    F2 = ctx.FormFunction_SNES(f["ucont"])
    assert np.array_equal(F1, F2)
    assert np.isfinite(F1).all()
    # boundary planes carry only 0.5*RHS_o - dP (momentum.c:1866-1938, 2322-2326)
    exp = 0.5 * f["rhs_o"][0] - f["dp"][0]
    assert np.allclose(F1[0], exp, rtol=0, atol=1e-15 * np.abs(exp).max())
    cs = ctx.download("CS")
    assert cs.min() >= 0 and cs.max() <= cfg["flags"]["max_cs"]
    ctx.close()


def test_async_download_equals_download(pkg):
    """vfs_download_async + vfs_download_wait deliver the same bytes as vfs_download, also with other work queued
    in between (the bench's e2e leg overlaps these copies with the next call's host->device copy)."""
    import torch
    capi, cases = pkg.capi, pkg.cases
    cfg = cases.scaled(cases.CONFIGS["c2_box256"], 40, 33, 37)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
    ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    f = cases.make_fields(cfg, met)
    for k, n in pc.FIELDS_IN:
        ctx.upload(n, f[k])
    ctx.Contra2Cart(); ctx.Compute_Smagorinsky_Constant_1(); ctx.Compute_eddy_viscosity_LES()
    cs = torch.empty((mz, my, mx), dtype=torch.float64).pin_memory()
    ucat = torch.empty((mz, my, mx, 3), dtype=torch.float64).pin_memory()
    ctx.download_async("CS", cs.data_ptr(), 0); ctx.download_async("UCAT", ucat.data_ptr(), 1)
    F = ctx.FormFunction_SNES(pc.krylov_x(f["ucont"]))      # queued behind the packs, overlapping the copies
    ctx.download_wait()
    assert np.array_equal(cs.numpy(), ctx.download("CS"))
    assert np.isfinite(F).all()
    # UCAT was packed BEFORE the residual recomputed it from the perturbed X: compare with a fresh run
    ctx2 = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
    ctx2.upload("COOR", cases.make_grid(cfg)); ctx2.FormMetrics()
    for k, n in pc.FIELDS_IN:
        ctx2.upload(n, f[k])
    ctx2.Contra2Cart()
    assert np.array_equal(ucat.numpy(), ctx2.download("UCAT"))
    ctx.close(); ctx2.close()


def test_multi_gpu_bitwise_equal_single_gpu():
    """k-slab decomposition over all visible GPUs (NCCL halos) == single-GPU result, bitwise."""
    import os, subprocess, sys, torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(root, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert "MULTIGPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["c2_small", "c3_small", "c1_small"])
def test_cuda_matches_golden_fixture(pkg, name):
    """Against the committed fixtures generated from the reference (no oracle/_ref needed)."""
    from test_cpu_golden import load_gold
    g, cfg = load_gold(name, pkg)
    err = pc.run_golden(cfg, g, device=0)
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_tiled_equals_staged(pkg):
    """The tiled / marching kernels against the one-thread-per-cell staged kernels.  Kernels that share
    arithmetic and summation order with the staged form must agree bitwise (Ucat, Ucont, metrics);
    the marching kernels that re-associate sums (separable test filters, fused residual) must agree
    to rounding, far inside the 1e-12 parity tolerance."""
    capi, cases = pkg.capi, pkg.cases
    for cfgname, dims in (("c2_box256", (70, 37, 45)), ("c3_turbine", (45, 30, 41)), ("c2_box256", (101, 67, 50))):
        cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        outs = []
        for fused in (0, 1):
            ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
            ctx.set_option(0, fused)
            ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
            for k, n in pc.FIELDS_IN:
                ctx.upload(n, f[k])
            outs.append(pc.run_path(ctx, pc.krylov_x(f["ucont"])))
            ctx.close()
        for n in outs[0]:
            if n in ("F", "CS", "NU_T", "FUSED_RHS", "FUSED_CS", "FUSED_NU_T"):
                assert pc.relerr(outs[1][n], outs[0][n]) <= 5e-13, (cfgname, n, pc.relerr(outs[1][n], outs[0][n]))
                if n in ("F", "FUSED_RHS"):       # mask-driven zeros are exact; Cs = max(C, 0) may flip at C ~ 0
                    assert np.array_equal(outs[0][n] == 0, outs[1][n] == 0), (cfgname, n)
            else:
                assert np.array_equal(outs[0][n], outs[1][n]), (cfgname, n)


@pytest.mark.parametrize("key", [12])
def test_fused_variants_bitwise_equal_staged_chain(pkg, key):
    """Option 12: Fp evaluated inside the projection kernel (ProjFpMarch) == FpCell + Fp planes + Project, bitwise
    (same arithmetic, same operand order), on a multi-tile grid with periodic i and k, and on the masked c3 case."""
    capi, cases = pkg.capi, pkg.cases
    for cfgname, dims in (("c2_box256", (101, 67, 50)), ("c3_turbine", (70, 37, 45))):
        cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        outs = []
        for val in (0, 1, 2):
            ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
            ctx.set_option(key, val)
            ctx.set_option(17, 0)          # the scalar FpCell: same arithmetic as the in-projection variants, instruction for instruction
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
            assert np.array_equal(outs[0][n], outs[1][n]), (cfgname, n, pc.relerr(outs[1][n], outs[0][n]))
            assert np.array_equal(outs[0][n], outs[2][n]), (cfgname, n, 'box', pc.relerr(outs[2][n], outs[0][n]))


def test_graph_replay_equals_eager(pkg):
    """vfs_rhs_les_fused replayed as a CUDA graph gives bitwise the same fields as eager launches."""
    capi, cases = pkg.capi, pkg.cases
    cfg = cases.scaled(cases.CONFIGS["c2_box256"], 70, 37, 45)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    res = []
    for graph in (0, 1):
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
        ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
        met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
        f = cases.make_fields(cfg, met)
        for k, n in pc.FIELDS_IN:
            ctx.upload(n, f[k])
        ctx.set_option(1, graph)
        for _ in range(4):          # eager, capture, replay, replay
            ctx.upload("UCONT", f["ucont"])
            ctx.rhs_les_fused()
        res.append({n: ctx.download(n) for n in ("RHS", "UCAT", "CS", "NU_T")})
        ctx.close()
    for n in res[0]:
        assert np.array_equal(res[0][n], res[1][n]), n


@pytest.mark.parametrize("key", [18, 21])
@pytest.mark.parametrize("name,dims", [("c2_box256", (70, 37, 45)), ("c3_turbine", (45, 30, 41))])
def test_unit_overlap_bitwise(pkg, name, dims, key):
    """Option 18 (the residual's Contra2Cart + IB_BC on a second stream beside LES pass 3 / nu_t inside vfs_rhs_les_fused)
    and option 21 (nu_t written by LES pass 3 inside the unit instead of a separate pass) change the schedule only:
    results bitwise equal with the option off, eagerly and as graph replays."""
    capi, cases = pkg.capi, pkg.cases
    cfg = cases.scaled(cases.CONFIGS[name], *dims)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    xyz = cases.make_grid(cfg)
    res = []
    for ovl in (0, 1):
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
        ctx.set_option(key, ovl)
        ctx.upload("COOR", xyz); ctx.FormMetrics()
        if ovl == 0:
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
        for k, n in pc.FIELDS_IN:
            ctx.upload(n, f[k])
        outs = []
        for _ in range(4):          # eager, capture, replay, replay
            ctx.upload("UCONT", f["ucont"])
            ctx.rhs_les_fused()
            outs.append({n: ctx.download(n) for n in ("RHS", "UCAT", "UCONT", "CS", "NU_T")})
        res.append(outs)
        ctx.close()
    for it in range(4):
        for n in res[0][it]:
            assert np.array_equal(res[0][it][n], res[1][it][n]), (it, n)


def test_full_size_kernel_families_agree(pkg):
    """BASELINE.json's bench size (256^3 nodes): the oracle cannot run there in seconds, so the marching /
    TMA kernels (what the bench times) are checked against the one-thread-per-cell staged kernels (the literal
    restatement, itself oracle-checked at small sizes) on the same device-resident case — every tile, k-chunk
    and the full k-march length of the benchmark configuration — plus two size-independent properties:
    Contra2Cart is idempotent, and the whole unit is deterministic (two runs agree bitwise)."""
    import torch
    if torch.cuda.mem_get_info(0)[0] < 40e9:
        pytest.skip("needs ~35 GB of free device memory")
    capi, cases = pkg.capi, pkg.cases
    cfg = dict(cases.CONFIGS["c2_box256"])
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ctxs = []
    for fused in (0, 1):
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
        ctx.set_option(0, fused)
        ctxs.append(ctx)
    ctxs[0].upload("COOR", cases.make_grid(cfg)); ctxs[0].FormMetrics()
    met = dict(csi=ctxs[0].download("CSI"), eta=ctxs[0].download("ETA"), zet=ctxs[0].download("ZET"), aj=ctxs[0].download("AJ"))
    f = cases.make_fields(cfg, met)
    ctxs[1].upload("COOR", cases.make_grid(cfg)); ctxs[1].FormMetrics()
    outs = []
    for ctx in ctxs:
        for k, n in pc.FIELDS_IN:
            ctx.upload(n, f[k])
        ctx.rhs_les_fused()
        outs.append({n: ctx.download(n) for n in ("RHS", "UCAT", "CS", "NU_T")})
    for n in ("RHS", "CS", "NU_T"):
        assert pc.relerr(outs[1][n], outs[0][n]) <= 5e-13, (n, pc.relerr(outs[1][n], outs[0][n]))
    assert np.array_equal(outs[0]["RHS"] == 0, outs[1]["RHS"] == 0)
    assert np.array_equal(outs[0]["UCAT"], outs[1]["UCAT"])
    assert np.isfinite(outs[1]["RHS"]).all() and np.abs(outs[1]["RHS"]).max() > 0
    ctx = ctxs[1]
    ctx.Contra2Cart()                      # idempotent on its own output
    assert np.array_equal(ctx.download("UCAT"), outs[1]["UCAT"])
    ctx.upload("UCONT", f["ucont"])        # deterministic: the same unit again
    ctx.rhs_les_fused()
    for n in ("RHS", "CS", "NU_T"):
        assert np.array_equal(ctx.download(n), outs[1][n]), n
    for c in ctxs:
        c.close()


def test_restart_files_feed_the_device_path(pkg, tmp_path):
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
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
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
