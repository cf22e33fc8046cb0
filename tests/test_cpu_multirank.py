"""World-size-2 (gloo, CPU) test of the k-slab halo layer: the two-rank result must be BITWISE
equal to the single-rank result (no reductions are involved).  Kernel bodies run through
the test-only host emulation; the halo logic (vfs-wind_b200/halo.py) is the product code."""
import os
import socket
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_common as pc
import emu_loader

from parity_common import FIELDS_IN, run_path  # noqa: E402


def _worker(rank, world, port, tmp, cfgname, dims, flags_extra, bctype):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = pc.load_package()
    lib = emu_loader.load(pkg.capi)
    capi, cases = pkg.capi, pkg.cases
    cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
    cfg["flags"] = dict(cfg["flags"], **flags_extra)
    if bctype:
        cfg["bctype"] = bctype
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    kofs, nzl = capi.slab_partition(mz, world)[rank]
    g = np.load(os.path.join(tmp, "global.npz"))
    p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], kofs=kofs, nzl=nzl, rank=rank, nranks=world)
    ctx = capi.VfsContext(p, lib=lib)
    halo = pkg.halo.TorchHalo(rank, world, periodic_k=bool(cfg["flags"].get("kk_periodic")), device="cpu")
    halo.attach(ctx)
    sl = slice(kofs, kofs + nzl)
    ctx.upload("COOR", g["xyz"][sl])
    ctx.FormMetrics()
    for k, n in FIELDS_IN:
        ctx.upload(n, g[k][sl])
    out = run_path(ctx, g["x"][sl])
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), nex=halo.nexchanges, **out)
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfgname,dims,extra,bctype,world", [
    ("c2_box256", (13, 11, 19), {}, None, 2),                    # kk periodic: rank 0 <-> rank 1 wrap
    ("c3_turbine", (17, 13, 21), {}, None, 2),                   # non-periodic k, IBM masks, F_eul
    ("c2_box256", (13, 11, 23), {}, None, 3),                    # a rank with two interior slab boundaries + the periodic seam
    ("c2_box256", (13, 11, 19), dict(skew=1, clark=1, levelset_weno=5), None, 2),      # the flux / LES variants: Adv1-3 and the gradient planes travel too
    ("c3_turbine", (17, 13, 21), dict(inviscid=1), None, 2),
])
def test_two_ranks_bitwise_equal_single_rank(pkg, refdrv, tmp_path, cfgname, dims, extra, bctype, world):
    capi, cases = pkg.capi, pkg.cases
    lib = emu_loader.load(capi)
    cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
    cfg["flags"] = dict(cfg["flags"], **extra)
    if bctype:
        cfg["bctype"] = bctype
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    xyz = cases.make_grid(cfg)
    ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]), lib=lib)
    ctx.upload("COOR", xyz)
    ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    f = cases.make_fields(cfg, met)
    for k, n in FIELDS_IN:
        ctx.upload(n, f[k])
    x = f["ucont"] * (1.0 + 1e-3 * np.sin(np.arange(f["ucont"].size).reshape(f["ucont"].shape)))
    single = run_path(ctx, x)
    ctx.close()
    np.savez(os.path.join(tmp_path, "global.npz"), xyz=xyz, x=x, **f)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path), cfgname, dims, extra, bctype), nprocs=world, join=True)
    parts = [np.load(os.path.join(tmp_path, "rank%d.npz" % r)) for r in range(world)]
    print("exchanges per rank:", [int(pp["nex"]) for pp in parts])
    assert int(parts[0]["nex"]) > 5
    for n in ("F", "UCAT", "CS", "NU_T", "UCONT", "CSI", "AJ", "FUSED_RHS", "FUSED_UCAT", "FUSED_CS", "FUSED_NU_T", "PROJ_P", "PROJ_PHI", "PROJ_UCONT"):
        multi = np.concatenate([pp[n] for pp in parts], axis=0)
        assert np.array_equal(multi, single[n]), n


def test_slab_partition(pkg):
    sp = pkg.capi.slab_partition
    assert sp(258, 1) == [(0, 258)]
    parts = sp(258, 4)
    assert sum(n for _, n in parts) == 258 and parts[0][0] == 0
    assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(3))
