"""On-disk formats either side of the hot path: the PETSc binary Vec files VFS-Wind checkpoints with
(`vfield%06d_0.dat` = Ucont, `ufield` = Ucat, `nvfield` = Nvert, `cs_` ... — Source/main.c:376-659 read, :684-935
write), the `xyz.dat` axis grid file (Source/init.c:258-338) and `bcs.dat` (Source/init.c:503-510).

PETSc (3.1, VecView on a binary viewer): big-endian int32 cookie 1211214 (VEC_FILE_COOKIE), big-endian int32 length,
then that many big-endian float64.  For a DA global Vec the order is the natural one, [k][j][i][dof] — exactly the
host layout `vfs_upload` takes — so a reference restart state can be loaded into a context directly:

    ucont = read_vec(path + "/vfield000100_0.dat").reshape(mz, my, mx, 3); ctx.upload("UCONT", ucont)
"""
import numpy as np

VEC_FILE_COOKIE = 1211214


def read_vec(path):
    with open(path, "rb") as f:
        hdr = np.frombuffer(f.read(8), dtype=">i4")
        if hdr.size != 2 or int(hdr[0]) != VEC_FILE_COOKIE:
            raise ValueError("%s: not a PETSc binary Vec (cookie %r)" % (path, hdr[:1]))
        n = int(hdr[1])
        a = np.frombuffer(f.read(8 * n), dtype=">f8")
        if a.size != n:
            raise ValueError("%s: truncated (%d of %d values)" % (path, a.size, n))
    return a.astype(np.float64)


def write_vec(path, a):
    a = np.ascontiguousarray(a, dtype=np.float64).ravel()
    with open(path, "wb") as f:
        np.array([VEC_FILE_COOKIE, a.size], dtype=">i4").tofile(f)
        a.astype(">f8").tofile(f)


def read_xyz_dat(path):
    """`xyz.dat`: first line IM JM KM, then IM + JM + KM lines whose first / second / third column is the x / y / z
    axis coordinate (Source/init.c:258-338).  Returns node coordinates (KM+1, JM+1, IM+1, 3), last index unused."""
    with open(path) as f:
        IM, JM, KM = [int(v) for v in f.readline().split()[:3]]
        rows = [[float(v) for v in f.readline().split()[:3]] for _ in range(IM + JM + KM)]
    X = np.array([r[0] for r in rows[:IM]]); Y = np.array([r[1] for r in rows[IM:IM + JM]]); Z = np.array([r[2] for r in rows[IM + JM:]])
    xyz = np.zeros((KM + 1, JM + 1, IM + 1, 3))
    z, y, x = np.meshgrid(Z, Y, X, indexing="ij")
    xyz[:KM, :JM, :IM, 0], xyz[:KM, :JM, :IM, 1], xyz[:KM, :JM, :IM, 2] = x, y, z
    return xyz


def read_bcs_dat(path):
    """`bcs.dat`: the six boundary types (i-low, i-high, j-low, j-high, k-low, k-high)."""
    with open(path) as f:
        return [int(v) for v in f.read().split()[:6]]
