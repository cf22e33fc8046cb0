"""GPU twin of test_cpu_glue_dropin.py: reference-named entry points (host glue) -> C ABI -> CUDA."""
import pytest
from test_cpu_glue_dropin import run_dropin, TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dims", [("c2_box256", (21, 17, 25)), ("c3_turbine", (29, 21, 25)), ("variants", (29, 21, 25))])
def test_glue_dropin_cuda(pkg, refdrv, name, dims):
    from test_cpu_glue_dropin import variant_cfg
    cfg = variant_cfg(pkg, dims) if name == "variants" else pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = run_dropin(refdrv, pkg, "libvfsglue_cuda.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


def test_glue_snes_solve_cuda(pkg, refdrv):
    from test_cpu_glue_dropin import run_glue_solver
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS["c3_turbine"], 29, 21, 25)
    err = run_glue_solver(refdrv, pkg, "libvfsglue_cuda.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= 1e-10)}
    assert not bad, bad


def test_glue_periodic_turbine_array_cuda(pkg, refdrv):
    from test_cpu_glue_dropin import run_periodic_turbines
    err = run_periodic_turbines(refdrv, pkg, "libvfsglue_cuda.so")
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("name,dims", [("c2_box256", (21, 17, 25)), ("c3_turbine", (29, 21, 25))])
def test_glue_two_time_steps_cuda(pkg, refdrv, name, dims):
    from test_cpu_glue_dropin import run_two_time_steps
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = run_two_time_steps(refdrv, pkg, "libvfsglue_cuda.so", cfg)
    bad = {k: v for k, v in err.items() if not (v <= 1e-11)}
    assert not bad, bad
