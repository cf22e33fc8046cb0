"""On-disk formats either side of the hot path: the PETSc binary Vec files VFS-Wind checkpoints with
(`vfield%06d_0.dat` = Ucont, `ufield` = Ucat, `nvfield` = Nvert, `cs_` ... — Source/main.c:376-659 read, :684-935
write), the `xyz.dat` axis grid file and the general curvilinear `grid.dat` (ASCII or `-binary 1`, Source/init.c:258-376) and
`bcs.dat` (Source/init.c:503-510).

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


def read_grid_dat(path, binary=False, cl=1.0, L_dim=1.0):
    """`grid.dat` (Source/init.c:264-376): the block count, then per block `IM JM KM` followed by every node's x, then every
    node's y, then every node's z, each in [k][j][i] order — ASCII numbers, or with `-binary 1` native int32 / float64.
    Coordinates are scaled by L_dim / cl as the reference does (:343-368).  Returns a list (one entry per block) of node
    coordinate arrays (KM+1, JM+1, IM+1, 3), last index of every direction unused — the layout `vfs_upload(VFS_COOR)` takes."""
    blocks = []
    if binary:
        raw = open(path, "rb").read()
        pos = 0

        def ints(n):
            nonlocal pos
            v = np.frombuffer(raw, dtype=np.int32, count=n, offset=pos); pos += 4 * n
            return [int(x) for x in v]

        def doubles(n):
            nonlocal pos
            v = np.frombuffer(raw, dtype=np.float64, count=n, offset=pos); pos += 8 * n
            return v
        nb = ints(1)[0]
    else:
        tok = open(path).read().split()
        pos = 0

        def ints(n):
            nonlocal pos
            v = [int(t) for t in tok[pos:pos + n]]; pos += n
            return v

        def doubles(n):
            nonlocal pos
            v = np.array(tok[pos:pos + n], dtype=np.float64); pos += n
            return v
        nb = ints(1)[0]
    for _ in range(nb):
        IM, JM, KM = ints(3)
        xyz = np.zeros((KM + 1, JM + 1, IM + 1, 3))
        for c in range(3):
            v = doubles(IM * JM * KM)
            if v.size != IM * JM * KM:
                raise ValueError("%s: truncated" % path)
            xyz[:KM, :JM, :IM, c] = v.reshape(KM, JM, IM) / cl * L_dim
        blocks.append(xyz)
    return blocks


def write_grid_dat(path, blocks, binary=False):
    """Inverse of read_grid_dat for node arrays (KM+1, JM+1, IM+1, 3) (or (KM, JM, IM, 3) without the unused last index)."""
    with open(path, "wb" if binary else "w") as f:
        if binary:
            np.array([len(blocks)], dtype=np.int32).tofile(f)
        else:
            f.write("%d\n" % len(blocks))
        for xyz in blocks:
            a = np.asarray(xyz, dtype=np.float64)
            KM, JM, IM = a.shape[0], a.shape[1], a.shape[2]
            if np.all(a[-1] == 0) and np.all(a[:, -1] == 0) and np.all(a[:, :, -1] == 0):      # padded (KM+1, JM+1, IM+1) form
                KM, JM, IM = KM - 1, JM - 1, IM - 1
            if binary:
                np.array([IM, JM, KM], dtype=np.int32).tofile(f)
            else:
                f.write("%d %d %d\n" % (IM, JM, KM))
            for c in range(3):
                v = np.ascontiguousarray(a[:KM, :JM, :IM, c]).ravel()
                if binary:
                    v.tofile(f)
                else:
                    f.write("\n".join("%.17e" % x for x in v) + "\n")


# restart files Ucont_Read loads (Source/main.c:376-430): name pattern -> (context field(s), dof)
RESTART_FILES = (("vfield", ("UCONT", "UCONT_O"), 3), ("ufield", ("UCAT", "UCAT_OLD"), 3), ("pfield", ("P",), 1), ("nvfield", ("NVERT",), 1),
                 ("cs_", ("CS",), 1))      # cs_: the LES restart of Cs (main.c:588-608, written at :870)


def restart_path(path, name, ti, block=0):
    return "%s/%s%06d_%1d.dat" % (path, name, ti, block)      # (cs_ carries its own underscore: cs_000100_0.dat)


def read_restart(path, ti, mx, my, mz, block=0):
    """The reference's restart state of time step `ti` (Ucont_Read, Source/main.c:376-430: Ucont from `vfield`, P from
    `pfield`, Nvert_o from `nvfield`, Ucat from `ufield`; Ucont_o <- Ucont and lUcat_old <- Ucat, :420-428) as a dict
    context-field name -> array in the layout `vfs_upload` takes.  Missing optional files are skipped."""
    import os
    out = {}
    for name, fields, dof in RESTART_FILES:
        fn = restart_path(path, name, ti, block)
        if not os.path.exists(fn):
            continue
        a = read_vec(fn)
        if a.size != mx * my * mz * dof:
            raise ValueError("%s: %d values, expected %d x %d x %d x %d" % (fn, a.size, mz, my, mx, dof))
        a = a.reshape((mz, my, mx, 3) if dof == 3 else (mz, my, mx))
        for f in fields:
            out[f] = a
    return out


def write_restart(path, ti, fields, block=0):
    """Inverse: fields = dict with any of UCONT, UCAT, P, NVERT, CS (as downloaded from a context)."""
    for name, flds, dof in RESTART_FILES:
        if flds[0] in fields:
            write_vec(restart_path(path, name, ti, block), fields[flds[0]])
