"""On-disk formats either side of the hot path (vfs-wind_b200/petsc_io.py)."""
import os
import struct
import numpy as np
import pytest


def test_petsc_binary_vec_layout_and_roundtrip(pkg, tmp_path):
    io = pkg.petsc_io
    a = np.arange(24, dtype=np.float64).reshape(2, 2, 2, 3) * 0.25 - 1.0
    p = os.path.join(tmp_path, "vfield000010_0.dat")
    io.write_vec(p, a)
    raw = open(p, "rb").read()
    # PETSc 3.1 VecView (binary): big-endian cookie 1211214, length, then big-endian doubles in natural [k][j][i][dof] order
    assert struct.unpack(">ii", raw[:8]) == (1211214, 24)
    assert struct.unpack(">d", raw[8:16])[0] == -1.0 and struct.unpack(">d", raw[-8:])[0] == a.ravel()[-1]
    assert np.array_equal(io.read_vec(p).reshape(a.shape), a)
    open(p, "wb").write(raw[:-8])
    with pytest.raises(ValueError):
        io.read_vec(p)


def test_xyz_dat_and_bcs_dat_readers(pkg, tmp_path):
    io, cases = pkg.petsc_io, pkg.cases
    X, Y, Z = [0.0, 0.5, 1.5], [0.0, 0.25], [0.0, 1.0, 2.0, 4.0]
    p = os.path.join(tmp_path, "xyz.dat")
    with open(p, "w") as f:
        f.write("3 2 4\n")
        for x in X: f.write("%.15e 0 0\n" % x)
        for y in Y: f.write("0 %.15e 0\n" % y)
        for z in Z: f.write("0 0 %.15e\n" % z)
    xyz = io.read_xyz_dat(p)
    assert xyz.shape == (5, 3, 4, 3)
    assert xyz[3, 1, 2].tolist() == [1.5, 0.25, 4.0] and not xyz[4].any() and not xyz[:, 2].any() and not xyz[:, :, 3].any()
    b = os.path.join(tmp_path, "bcs.dat")
    open(b, "w").write("100 100 1 10 100 100\n")
    assert io.read_bcs_dat(b) == [100, 100, 1, 10, 100, 100]
    shipped = "/root/reference/Instructional_Cases/Test_10_ChannelFlow_Retau3000/xyz.dat"
    if os.path.exists(shipped):       # the C1 grid generator reproduces the shipped file
        ref = io.read_xyz_dat(shipped)
        assert np.abs(ref - cases.make_grid(cases.CONFIGS["c1_test10"])).max() < 1e-14


def test_grid_dat_ascii_and_binary_roundtrip(pkg, tmp_path):
    """`grid.dat` (init.c:264-376), ASCII and `-binary 1`: block count, IM JM KM, all x / all y / all z in [k][j][i]
    order.  A curvilinear two-block file written and read back bit-exactly; the first block in the layout FormMetrics takes."""
    io, cases = pkg.petsc_io, pkg.cases
    a = cases.make_grid(cases.scaled(cases.CONFIGS["c3_turbine"], 11, 8, 9))
    b = cases.make_grid(cases.scaled(cases.CONFIGS["c2_box256"], 7, 9, 6))
    for binary in (False, True):
        path = str(tmp_path / ("grid_%d.dat" % binary))
        io.write_grid_dat(path, [a, b], binary=binary)
        back = io.read_grid_dat(path, binary=binary)
        assert len(back) == 2
        assert np.array_equal(back[0], a) and np.array_equal(back[1], b)
        scaled = io.read_grid_dat(path, binary=binary, cl=2.0, L_dim=3.0)
        assert np.array_equal(scaled[0], a / 2.0 * 3.0)
    # the ASCII file starts exactly as the reference's reader expects: "<blocks>\n<IM> <JM> <KM>\n<x of node (0,0,0)>"
    head = open(str(tmp_path / "grid_0.dat")).read().split()[:5]
    assert head[:4] == ["2", "11", "8", "9"] and float(head[4]) == a[0, 0, 0, 0]


def test_restart_state_roundtrip(pkg, tmp_path):
    """Ucont_Read's file set (main.c:376-430): vfield / ufield / pfield / nvfield of a time step, written from arrays in
    the upload layout and read back as the dict of context fields (Ucont_o and Ucat_old aliased as the reference does)."""
    io = pkg.petsc_io
    rng = np.random.default_rng(3)
    mx, my, mz = 7, 6, 5
    f = dict(UCONT=rng.standard_normal((mz, my, mx, 3)), UCAT=rng.standard_normal((mz, my, mx, 3)), P=rng.standard_normal((mz, my, mx)),
             NVERT=(rng.random((mz, my, mx)) > 0.8) * 3.0, CS=rng.random((mz, my, mx)) * 0.1)
    io.write_restart(str(tmp_path), 120, f)
    assert os.path.exists(os.path.join(tmp_path, "cs_000120_0.dat"))
    assert os.path.exists(os.path.join(tmp_path, "vfield000120_0.dat")) and os.path.exists(os.path.join(tmp_path, "nvfield000120_0.dat"))
    back = io.read_restart(str(tmp_path), 120, mx, my, mz)
    for k in ("UCONT", "UCAT", "P", "NVERT", "CS"):
        assert np.array_equal(back[k], f[k]), k
    assert np.array_equal(back["UCONT_O"], f["UCONT"]) and np.array_equal(back["UCAT_OLD"], f["UCAT"])
    with pytest.raises(ValueError):
        io.read_restart(str(tmp_path), 120, mx + 1, my, mz)
