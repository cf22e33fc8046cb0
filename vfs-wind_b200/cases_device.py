"""The seeded synthetic cases of cases.py generated ON THE GPU (torch), for k-slabs too large to build on the
host: the weak-scaling grid of BASELINE.json configs[4] is 2048 x 1024 x 64 nodes per GPU (134 M nodes, 3.2 GB
per vector field), 8 of them per box.  Same recipe as cases.make_grid / make_masks / make_fields (SURVEY 8d:
stretched curvilinear grid, smoothed noisy bulk profile, fluxes consistent with the metrics, IBM masks and
actuator forcing placed in GLOBAL coordinates) with torch's generator instead of numpy's, so the numbers differ
from the host generator's — it feeds bench.py only; parity tests use cases.py.  Arrays are handed to the library
as device pointers (vfs_upload accepts them)."""
import numpy as np


def _t():
    import torch
    return torch


def _smooth121(a, axes=(0, 1, 2)):
    torch = _t()
    for ax in axes:
        a = 0.25 * torch.roll(a, 1, ax) + 0.5 * a + 0.25 * torch.roll(a, -1, ax)
    return a


def make_grid_device(cfg, kofs, nzl, dev):
    """cases.make_grid for planes [kofs, kofs + nzl) as a device tensor (nzl, my, mx, 3)."""
    torch = _t()
    IM, JM, KM = cfg["IM"], cfg["JM"], cfg["KM"]
    mx, my = IM + 1, JM + 1
    f64 = torch.float64
    xi = np.arange(IM) / (IM - 1.0)
    et = np.arange(JM) / (JM - 1.0)
    ksel = np.arange(kofs, min(kofs + nzl, KM))
    if cfg["grid"] == "test10":
        X, Y, Z = -0.6 + 1.2 * xi, 0.4 * et, 2.0 * (np.arange(KM) / (KM - 1.0))
        warp = False
    else:
        Lx, H = 2.0, 1.0
        Lz = 3.0 * (KM + 1) / 256.0 if cfg.get("weak_k") else 3.0
        beta = 2.0
        X = Lx * xi
        Y = H * (1.0 + np.tanh(beta * (et - 1.0)) / np.tanh(beta))
        r = 1.0 + 0.3 * np.sin(2 * np.pi * (np.arange(KM) + 0.5) / KM)
        Z = Lz * np.concatenate([[0.0], np.cumsum(0.5 * (r[1:] + r[:-1]))]) / np.sum(0.5 * (r[1:] + r[:-1]))
        warp = True
    n = len(ksel)
    x = torch.as_tensor(X, device=dev, dtype=f64)[None, None, :].expand(n, JM, IM)
    y = torch.as_tensor(Y, device=dev, dtype=f64)[None, :, None].expand(n, JM, IM)
    z = torch.as_tensor(Z[ksel], device=dev, dtype=f64)[:, None, None].expand(n, JM, IM)
    if warp:
        a = 0.02 * Lx
        x = x + a * torch.sin(2 * np.pi * y / H) * torch.sin(2 * np.pi * z / Lz)
        y = y + 0.01 * H * torch.sin(2 * np.pi * x / Lx) * torch.sin(np.pi * y / H) * torch.cos(2 * np.pi * z / Lz)
        z = z + 0.01 * Lz * torch.sin(2 * np.pi * x / Lx) * torch.sin(np.pi * y / H)
    xyz = torch.zeros((nzl, my, mx, 3), device=dev, dtype=f64)
    xyz[:n, :JM, :IM, 0] = x
    xyz[:n, :JM, :IM, 1] = y
    xyz[:n, :JM, :IM, 2] = z
    return xyz


def make_masks_device(cfg, kofs, nzl, mz, dev):
    torch = _t()
    mx, my = cfg["IM"] + 1, cfg["JM"] + 1
    nv = torch.zeros((nzl, my, mx), device=dev, dtype=torch.float64)
    if not cfg.get("masks"):
        return nv
    k = (torch.arange(nzl, device=dev) + kofs)[:, None, None]
    j = torch.arange(my, device=dev)[None, :, None]
    i = torch.arange(mx, device=dev)[None, None, :]
    solid = torch.zeros((nzl, my, mx), device=dev, dtype=torch.bool)
    nrow, ncol = (4, 8) if mx > 600 else (1, 2)
    hub = max(4, int(0.35 * my))
    tw = max(1, mx // 128)
    for r in range(nrow):
        for c in range(ncol):
            ci = int(mx * (c + 0.5) / ncol)
            ck = int(mz * (0.25 + 0.5 * (r + 0.5) / nrow))
            solid |= ((i - ci).abs() <= tw) & ((k - ck).abs() <= tw) & (j >= 1) & (j <= hub)
            solid |= (((i - ci) / (2.0 * tw + 1)) ** 2 + ((j - hub) / (1.5 * tw + 1)) ** 2 + ((k - ck) / (3.0 * tw + 2)) ** 2) <= 1.0
    edge = ((k == 0) | (k == mz - 1) | (j == 0) | (j == my - 1) | (i == 0) | (i == mx - 1)).expand(nzl, my, mx)
    solid &= ~edge
    near = torch.zeros_like(solid)
    for ax in range(3):      # (slab ends: the neighbouring slab's bodies are not seen; bodies sit well inside the slabs' k ranges or are cut identically)
        near |= torch.roll(solid, 1, ax) | torch.roll(solid, -1, ax)
    nv[solid] = 3.0
    nv[near & ~solid & ~edge] = 1.0
    return nv


def fill_context(ctx, cfg, kofs, nzl, rank, device):
    """Grid -> FormMetrics on the device -> fields, all without leaving HBM.  Returns a small dict (no host arrays)."""
    torch = _t()
    dev = torch.device("cuda", device)
    f64 = torch.float64
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    shp = (nzl, my, mx)

    def up(name, t):
        t = t.contiguous()
        torch.cuda.synchronize(dev)
        ctx.upload_ptr(name, t.data_ptr())

    def down(name, dof):
        t = torch.empty(shp + ((3,) if dof == 3 else ()), device=dev, dtype=f64)
        ctx.download_ptr(name, t.data_ptr())
        return t
    xyz = make_grid_device(cfg, kofs, nzl, dev)
    up("COOR", xyz)
    del xyz
    ctx.FormMetrics()
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(cfg["seed"]) + 7919 * rank)

    def noise(shape):
        return torch.rand(shape, generator=gen, device=dev, dtype=f64) * 2.0 - 1.0
    eta = (torch.arange(my, device=dev, dtype=f64) - 0.5) / (my - 2.0)
    prof = 1.2 * (1.0 - (2 * eta.clamp(0, 1) - 1.0).abs() ** 8)
    w = prof[None, :, None] * (1.0 + 0.3 * noise(shp))
    ucat = torch.empty(shp + (3,), device=dev, dtype=f64)
    ucat[..., 0] = _smooth121(0.3 * noise(shp) * w)
    ucat[..., 1] = _smooth121(0.3 * noise(shp) * w)
    ucat[..., 2] = _smooth121(w)
    del w
    nvert = make_masks_device(cfg, kofs, nzl, mz, dev)
    ucat[nvert > 0.1] = 0.0
    up("NVERT", nvert)
    up("UCAT", ucat)
    up("UCAT_OLD", ucat + 1e-3 * _smooth121(noise(shp + (3,))))

    def face(a, ax):
        return 0.5 * (a + torch.roll(a, -1, ax))
    ucont = torch.zeros(shp + (3,), device=dev, dtype=f64)
    for comp, (name, ax) in enumerate((("CSI", 2), ("ETA", 1), ("ZET", 0))):
        met = down(name, 3)
        ucont[..., comp] = (face(ucat, ax) * face(met, ax)).sum(-1)
        if comp == 2:
            zet = met
        else:
            del met
    bc = cfg["bctype"]
    if bc[2] in (1, 10, -1, -2):
        ucont[:, 0, :, 1] = 0.0
    if bc[3] in (1, 10, -1, -2):
        ucont[:, my - 2, :, 1] = 0.0
    scale = float(ucont.abs().max())
    up("UCONT", ucont)
    up("UCONT_O", ucont + 1e-3 * scale * _smooth121(noise(shp + (3,))))
    # (Ucont_rm1 is read by the BDF2 assembly only, which this path never takes: not uploaded, its scalars stay unallocated)
    del ucont
    up("DP", 0.05 * scale * _smooth121(noise(shp + (3,))))
    up("RHS_O", 0.05 * scale * _smooth121(noise(shp + (3,))))
    f_eul = torch.zeros(shp + (3,), device=dev, dtype=f64)
    if cfg.get("forcing"):
        aj = down("AJ", 1)
        k = (torch.arange(nzl, device=dev) + kofs)[:, None, None]
        j = torch.arange(my, device=dev)[None, :, None]
        i = torch.arange(mx, device=dev)[None, None, :]
        nrow, ncol = (4, 8) if mx > 600 else (1, 2)
        hub = max(4, int(0.35 * my))
        R = max(3.0, 0.2 * my)
        fz = torch.zeros(shp, device=dev, dtype=f64)
        for r in range(nrow):
            for c in range(ncol):
                ci = int(mx * (c + 0.5) / ncol)
                ck = int(mz * (0.25 + 0.5 * (r + 0.5) / nrow)) - max(3, mz // 32)
                rad = torch.sqrt((i - ci) ** 2.0 + (j - hub) ** 2.0)
                dd = (k - ck).abs() / 2.0
                delta = torch.where(dd < 1.0, 0.5 * (1.0 + torch.cos(np.pi * dd)) / 2.0, torch.zeros((), device=dev, dtype=f64))
                fz = fz + (-0.4) * delta * (rad <= R)
        f_eul[..., 2] = fz * face(zet, 0)[..., 2] * face(aj, 0)
        del aj, fz
    del zet
    up("F_EUL", f_eul)
    del f_eul, ucat, nvert
    torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    return {"generated": "device", "scale": scale}
