"""CPU twin of the Newton-Krylov parity test: the solver logic (vfs_solver.h compiled for the host, -DVFS_EMU,
test-only) on the emulated residual against the numpy restatement driving the oracle residual."""
import pytest
import emu_loader
import solver_common as sc


@pytest.fixture(scope="module")
def emu(pkg):
    return emu_loader.load(pkg.capi)


@pytest.mark.parametrize("name,dims,kw", [
    ("c2_box256", (13, 11, 15), {}),
    ("c3_turbine", (17, 13, 15), {}),
    ("c2_box256", (13, 11, 15), dict(restart=3, use_ew=0, ksp_rtol=1e-9, rtol=1e-10)),          # several GMRES restart cycles
    ("c3_turbine", (15, 11, 13), dict(trust_region=0, use_ew=0)),
])
def test_emulated_momentum_solve_matches_host_restatement(pkg, refdrv, emu, name, dims, kw):
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    sc.check(*sc.run_solver_parity(cfg, refdrv, lib=emu, **kw))


@pytest.mark.parametrize("flags", [dict(skew=1, clark=1), dict(levelset_weno=5), dict(i_homo_filter=1, k_homo_filter=1)])
def test_emulated_momentum_solve_with_flux_variants(pkg, refdrv, emu, flags):
    """The solver drives whatever residual the flags select: the skew-symmetric / Clark / WENO3 fluxes and the homogeneous
    Cs averaging included (one-thread-per-face kernels, momentum.c:754-923)."""
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 15, 11, 13)
    cfg["flags"] = dict(cfg["flags"], **flags)
    sc.check(*sc.run_solver_parity(cfg, refdrv, lib=emu))
