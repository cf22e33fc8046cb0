"""torchrun script: N-GPU k-slab run (NCCL halos) must be BITWISE equal to the 1-GPU run.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402
from parity_common import FIELDS_IN, run_path  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    pkg = pc.load_package()
    capi, cases = pkg.capi, pkg.cases
    ok = True
    for cfgname, dims in (("c2_box256", (61, 45, 16 * world + 7)), ("c3_turbine", (53, 37, 12 * world + 9))):
        cfg = cases.scaled(cases.CONFIGS[cfgname], *dims)
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        xyz = cases.make_grid(cfg)
        # single-GPU result, computed redundantly on every rank's own device
        ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], device=lrank))
        ctx.upload("COOR", xyz); ctx.FormMetrics()
        met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
        f = cases.make_fields(cfg, met)
        for k, n in FIELDS_IN:
            ctx.upload(n, f[k])
        x = f["ucont"] * (1.0 + 1e-3 * np.sin(np.arange(f["ucont"].size).reshape(f["ucont"].shape)))
        single = run_path(ctx, x)
        ctx.close()
        kofs, nzl = capi.slab_partition(mz, world)[rank]
        p = capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"], kofs=kofs, nzl=nzl, rank=rank, nranks=world, device=lrank)
        ctx = capi.VfsContext(p)
        if os.environ.get("VFS_HALO") == "torch":      # the torch.distributed callback layer
            halo = pkg.halo.TorchHalo(rank, world, periodic_k=bool(cfg["flags"].get("kk_periodic")), device=torch.device("cuda", lrank))
            halo.attach(ctx)
        else:                                           # in-library NCCL halo layer (default)
            ctx.nccl_init(dist, device=torch.device("cuda", lrank))
        sl = slice(kofs, kofs + nzl)
        ctx.upload("COOR", xyz[sl]); ctx.FormMetrics()
        for k, n in FIELDS_IN:
            ctx.upload(n, f[k][sl])
        out = run_path(ctx, x[sl])
        for n in ("F", "UCAT", "CS", "NU_T", "UCONT", "CSI", "AJ"):
            same = np.array_equal(out[n], single[n][sl])
            ok = ok and same
            if not same:
                print("rank %d %s %s MISMATCH max %.3e" % (rank, cfgname, n, np.abs(out[n] - single[n][sl]).max()), flush=True)
        ctx.close()
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK %s world=%d" % ("PASS" if t.item() == 1.0 else "FAIL", world), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
