"""Seeded synthetic cases for the momentum RHS + LES path.

Everything here is host-side NumPy producing flat `[k][j][i][c]` FP64 arrays in the reference's DA
"global vector" layout (mx = IM+1, my = JM+1, mz = KM+1 nodes; index 0 and m-1 are ghost cells,
Source/init.c:157-160).  Used by bench.py and tests; no CUDA, no oracle imports.

Velocity recipe follows the reference initial condition (Source/bcs.c:3177-3208): a bulk profile w
perturbed by uniform noise, u = 0.1*n3*w, v = 0.1*n2*w, then smoothed once with a (1-2-1)^3 filter.
"""
import numpy as np

# bcs.dat / control.dat of the instructional cases the configs derive from
# (Instructional_Cases/Test_10_ChannelFlow_Retau3000/{control.dat,bcs.dat};
#  Instructional_Cases/Test_09_ModelWindTurbine/TSR4/{control.dat,bcs.dat})
CONFIGS = {
    # C1: Test_10 channel, shipped grid 121x41x61 (uniform in x,z; generated identically here)
    "c1_test10": dict(IM=121, JM=41, KM=61, grid="test10", seed=101, ren=62500.0, dt=1e-3, bctype=[100, 100, 1, 10, 100, 100],
                      flags=dict(les=2, laplacian=1, second_order=1, ii_periodic=1, kk_periodic=1, viscosity_wallmodel=1, max_cs=0.2)),
    # C2: synthetic stretched curvilinear box 256^3 with dynamic Smagorinsky (bench workload)
    "c2_box256": dict(IM=255, JM=255, KM=255, grid="stretched", seed=202, ren=62500.0, dt=1e-3, bctype=[100, 100, 1, 10, 100, 100],
                      flags=dict(les=2, ii_periodic=1, kk_periodic=1, max_cs=0.2)),
    # C3: Test_09-style turbine grid with IBM masks and actuator forcing, 512x256x256
    "c3_turbine": dict(IM=511, JM=255, KM=255, grid="stretched", seed=303, ren=429407.0, dt=1e-3, bctype=[100, 100, 1, 10, 5, 4],
                       flags=dict(les=2, ii_periodic=1, immersed=1, rotor_model=1, max_cs=0.1), masks=True, forcing=True),
    # C4: wind-farm ABL LES 1024x512x256 with turbine rows (k-slab sharded)
    "c4_farm": dict(IM=1023, JM=511, KM=255, grid="stretched", seed=404, ren=429407.0, dt=1e-3, bctype=[100, 100, 1, 10, 5, 4],
                    flags=dict(les=2, ii_periodic=1, immersed=1, rotor_model=1, max_cs=0.1), masks=True, forcing=True),
    # C5: weak-scaling grid 2048x1024x512
    "c5_weak": dict(IM=2047, JM=1023, KM=511, grid="stretched", seed=505, ren=62500.0, dt=1e-3, bctype=[100, 100, 1, 10, 100, 100],
                    flags=dict(les=2, ii_periodic=1, kk_periodic=1, max_cs=0.2)),
}


def scaled(cfg, IM, JM, KM):
    """Same physics/flags as `cfg` on a smaller grid (parity-test sizes)."""
    c = dict(cfg)
    c.update(IM=IM, JM=JM, KM=KM)
    return c


def make_grid(cfg, kofs=0, nzl=None):
    """Node coordinates, shape (nzl, my, mx, 3) for the k-planes [kofs, kofs+nzl) (default: all);
    the last index in each direction is unused (zero), exactly as the reference leaves it
    (Source/init.c:340-376 fills only IM*JM*KM nodes)."""
    IM, JM, KM = cfg["IM"], cfg["JM"], cfg["KM"]
    mx, my, mz = IM + 1, JM + 1, KM + 1
    if nzl is None:
        nzl = mz - kofs
    xi = np.arange(IM) / (IM - 1.0)
    et = np.arange(JM) / (JM - 1.0)
    ze = np.arange(KM) / (KM - 1.0)
    ksel = np.arange(kofs, min(kofs + nzl, KM))
    if cfg["grid"] == "test10":
        # shipped xyz.dat: x in [-0.6,0.6], y in [0,0.4], z in [0,2], all uniform (file values agree to 3e-15)
        X = -0.6 + 1.2 * xi
        Y = 0.4 * et
        Z = 2.0 * ze
        x, y, z = np.meshgrid(X, Y, Z[ksel], indexing="ij")
    else:
        Lx, H = 2.0, 1.0
        Lz = 3.0 * (KM + 1) / 256.0 if cfg.get("weak_k") else 3.0
        beta = 2.0
        X = Lx * xi
        Y = H * (1.0 + np.tanh(beta * (et - 1.0)) / np.tanh(beta))        # wall-clustered at y=0
        r = 1.0 + 0.3 * np.sin(2 * np.pi * (np.arange(KM) + 0.5) / KM)      # smooth, periodic stretching in z
        Z = Lz * np.concatenate([[0.0], np.cumsum(0.5 * (r[1:] + r[:-1]))]) / np.sum(0.5 * (r[1:] + r[:-1]))
        x, y, z = np.meshgrid(X, Y, Z[ksel], indexing="ij")
        a = 0.02 * Lx
        # smooth warp so that all nine metric components are non-zero; periodic in x and z
        x = x + a * np.sin(2 * np.pi * y / H) * np.sin(2 * np.pi * z / Lz)
        y = y + 0.01 * H * np.sin(2 * np.pi * x / Lx) * np.sin(np.pi * y / H) * np.cos(2 * np.pi * z / Lz)
        z = z + 0.01 * Lz * np.sin(2 * np.pi * x / Lx) * np.sin(np.pi * y / H)
    xyz = np.zeros((nzl, my, mx, 3))
    n = len(ksel)
    xyz[:n, :JM, :IM, 0] = x.transpose(2, 1, 0)
    xyz[:n, :JM, :IM, 1] = y.transpose(2, 1, 0)
    xyz[:n, :JM, :IM, 2] = z.transpose(2, 1, 0)
    return xyz


def make_grid_slab(cfg, kofs, nzl):
    return make_grid(cfg, kofs, nzl)


def _smooth121(a, axes=(0, 1, 2)):
    for ax in axes:
        a = 0.25 * np.roll(a, 1, ax) + 0.5 * a + 0.25 * np.roll(a, -1, ax)
    return a


def make_masks(cfg, rng):
    """Nvert: 3 inside tower/nacelle boxes and ellipsoids, 1 on their one-cell fluid-side shell,
    0 elsewhere (integers stored as doubles, Source/ibm.c:231-232,518-519)."""
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    nzl = mz
    nv = np.zeros((nzl, my, mx))
    if not cfg.get("masks"):
        return nv
    kofs = 0
    if cfg.get("mask_window"):          # k-slab of a larger grid: bodies are placed in GLOBAL k (kofs, global mz)
        kofs, mz = cfg["mask_window"]
    k, j, i = np.meshgrid(np.arange(nzl) + kofs, np.arange(my), np.arange(mx), indexing="ij")
    solid = np.zeros((nzl, my, mx), bool)
    nrow = 4 if mx > 600 else 1
    ncol = 8 if mx > 600 else 2
    for r in range(nrow):
        for c in range(ncol):
            ci = int(mx * (c + 0.5) / ncol)
            ck = int(mz * (0.25 + 0.5 * (r + 0.5) / nrow))
            hub = max(4, int(0.35 * my))
            tw = max(1, mx // 128)
            solid |= (abs(i - ci) <= tw) & (abs(k - ck) <= tw) & (j >= 1) & (j <= hub)           # tower (box)
            solid |= (((i - ci) / (2.0 * tw + 1)) ** 2 + ((j - hub) / (1.5 * tw + 1)) ** 2 + ((k - ck) / (3.0 * tw + 2)) ** 2) <= 1.0  # nacelle (ellipsoid)
    kend = (k == 0) | (k == mz - 1)
    solid[:, 0, :] = solid[:, -1, :] = False
    solid[kend] = False
    solid[:, :, 0] = solid[:, :, -1] = False
    near = np.zeros_like(solid)
    for ax in range(3):
        near |= np.roll(solid, 1, ax) | np.roll(solid, -1, ax)
    nv[solid] = 3.0
    shell = near & ~solid
    shell[:, 0, :] = shell[:, -1, :] = False
    shell[kend] = False
    shell[:, :, 0] = shell[:, :, -1] = False
    nv[shell] = 1.0
    return nv


def make_fields(cfg, metrics):
    """State vectors for one RHS+LES evaluation.  `metrics` = dict(csi, eta, zet, aj) as owned
    (mz,my,mx[,3]) arrays (from FormMetrics — device or oracle — so the fluxes are consistent with
    the geometry).  Returns dict of (mz,my,mx[,3]) arrays: ucat, ucont, ucont_o, ucont_rm1, ucat_old,
    dp, f_eul, nvert, rhs_o (zeros; callers overwrite with Formfunction_2(Ucont_o) if wanted)."""
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    rng = np.random.default_rng(cfg["seed"])
    eta = (np.arange(my) - 0.5) / (my - 2.0)
    prof = 1.2 * (1.0 - np.abs(2 * np.clip(eta, 0, 1) - 1.0) ** 8)
    w = prof[None, :, None] * (1.0 + 0.3 * rng.uniform(-1, 1, (mz, my, mx)))
    u = 0.3 * rng.uniform(-1, 1, (mz, my, mx)) * w
    v = 0.3 * rng.uniform(-1, 1, (mz, my, mx)) * w
    ucat = np.stack([_smooth121(u), _smooth121(v), _smooth121(w)], -1)
    nvert = make_masks(cfg, rng)
    ucat[nvert > 0.1] = 0.0
    csi, et, zet = metrics["csi"], metrics["eta"], metrics["zet"]
    ucont = np.zeros((mz, my, mx, 3))

    def face(a, ax):
        return 0.5 * (a + np.roll(a, -1, ax))
    ucont[..., 0] = np.sum(face(ucat, 2) * face(csi, 2), -1)
    ucont[..., 1] = np.sum(face(ucat, 1) * face(et, 1), -1)
    ucont[..., 2] = np.sum(face(ucat, 0) * face(zet, 0), -1)
    bc = cfg["bctype"]
    if bc[2] in (1, 10, -1, -2):
        ucont[:, 0, :, 1] = 0.0
    if bc[3] in (1, 10, -1, -2):
        ucont[:, my - 2, :, 1] = 0.0
    scale = np.abs(ucont).max()
    out = dict(ucat=ucat, ucont=ucont, nvert=nvert)
    out["ucont_o"] = ucont + 1e-3 * scale * _smooth121(rng.uniform(-1, 1, ucont.shape), (0, 1, 2))
    out["ucont_rm1"] = ucont + 2e-3 * scale * _smooth121(rng.uniform(-1, 1, ucont.shape), (0, 1, 2))
    out["ucat_old"] = ucat + 1e-3 * _smooth121(rng.uniform(-1, 1, ucat.shape), (0, 1, 2))
    out["dp"] = 0.05 * scale * _smooth121(rng.uniform(-1, 1, ucont.shape), (0, 1, 2))
    out["rhs_o"] = 0.05 * scale * _smooth121(rng.uniform(-1, 1, ucont.shape), (0, 1, 2))
    f_eul = np.zeros((mz, my, mx, 3))
    if cfg.get("forcing"):
        # actuator-disk style forcing: smoothed 2h delta (Source/rotor_model.c:5130 dfunc_2h) around
        # rotor planes at hub height, projected on the face area vectors like Calc_F_eul (:3785-3829)
        kofs, mzg = cfg.get("mask_window") or (0, mz)
        k, j, i = np.meshgrid(np.arange(mz) + kofs, np.arange(my), np.arange(mx), indexing="ij")
        nrow = 4 if mx > 600 else 1
        ncol = 8 if mx > 600 else 2
        hub = max(4, int(0.35 * my))
        R = max(3.0, 0.2 * my)
        for r in range(nrow):
            for c in range(ncol):
                ci = int(mx * (c + 0.5) / ncol)
                ck = int(mzg * (0.25 + 0.5 * (r + 0.5) / nrow)) - max(3, mzg // 32)
                rad = np.sqrt((i - ci) ** 2.0 + (j - hub) ** 2.0)
                d = np.abs(k - ck) / 2.0
                delta = np.where(d < 1.0, 0.5 * (1.0 + np.cos(np.pi * d)) / 2.0, 0.0)
                fz = -0.4 * delta * (rad <= R)
                f_eul[..., 0] += fz * face(zet, 2)[..., 2] * 0.0
                f_eul[..., 2] += fz * face(zet, 0)[..., 2] * face(metrics["aj"], 0)
    out["f_eul"] = f_eul
    # pressure: smooth random field, zero inside bodies (input of Pressure_Gradient, momentum.c:203)
    pr = 0.5 * _smooth121(rng.uniform(-1, 1, (mz, my, mx)))
    pr[nvert > 1.1] = 0.0
    out["p"] = pr
    return out
