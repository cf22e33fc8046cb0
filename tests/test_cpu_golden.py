"""Committed golden vectors (generated from the reference by tests/golden/make_golden.py):
 * the fixtures still match what oracle/_ref produces today (pins the oracle build);
 * the kernel logic (host emulation, test-only) matches the fixtures without oracle/_ref."""
import os
import numpy as np
import pytest
import parity_common as pc
import emu_loader

GOLD = os.path.join(pc.ROOT, "tests", "golden")
TOL = 1e-12


def load_gold(name, pkg):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[str(g["cfgname"])], *[int(v) for v in g["dims"]])
    return g, cfg


@pytest.mark.parametrize("name", ["c2_small", "c3_small", "c1_small"])
def test_emulated_kernels_match_golden(pkg, name):
    g, cfg = load_gold(name, pkg)
    err = pc.run_golden(cfg, g, lib=emu_loader.load(pkg.capi))
    bad = {k: v for k, v in err.items() if not (v <= TOL)}
    assert not bad, bad


@pytest.mark.parametrize("name", ["c2_small", "c3_small", "c1_small"])
def test_reference_build_reproduces_golden(pkg, refdrv, name):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g, cfg = load_gold(name, pkg)
    out = mg.reference_outputs(cfg)
    for k, v in out.items():
        assert np.array_equal(v, g[k]), k


def test_c_abi_exports_every_declared_symbol(pkg):
    """include/vfs_b200.h <-> libvfs_b200.so: every declared entry point is exported (no compute
    calls: there is no GPU here)."""
    import ctypes, re
    hdr = open(os.path.join(pc.ROOT, "include", "vfs_b200.h")).read()
    declared = set(re.findall(r"\b(vfs_[a-z0-9_]+)\s*\(", hdr)) - {"vfs_halo_fn"}
    assert set(pkg.capi.EXPORTS) <= declared
    if not os.path.exists(pkg.capi.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = ctypes.CDLL(pkg.capi.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
