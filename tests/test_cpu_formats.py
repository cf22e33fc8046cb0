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
