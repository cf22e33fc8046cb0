"""torchrun script: N-GPU k-slab run (NCCL halos) must be BITWISE equal to the 1-GPU run.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py
(the comparison itself is vfs-wind_b200/selfcheck.py, which bench.py also runs before timing N > 1)"""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
    pkg = pc.load_package()

    def make_halo(ctx, cfg):
        if os.environ.get("VFS_HALO") == "torch":      # the torch.distributed callback layer
            pkg.halo.TorchHalo(rank, world, periodic_k=bool(cfg["flags"].get("kk_periodic")), device=torch.device("cuda", lrank)).attach(ctx)
        else:                                           # in-library NCCL halo layer (default)
            ctx.nccl_init(dist, device=torch.device("cuda", lrank))
    ok = pkg.selfcheck.nrank_equals_1rank(pkg.capi, pkg.cases, rank, world, lrank, make_halo)
    ok = pkg.selfcheck.nrank_equals_1rank(pkg.capi, pkg.cases, rank, world, lrank, make_halo, pkg.selfcheck.variant_cases(world)) and ok
    if os.environ.get("VFS_HALO") != "torch":       # homogeneous Cs averaging: ncclAllReduce of the plane sums, 1e-12
        ok = pkg.selfcheck.nrank_equals_1rank(pkg.capi, pkg.cases, rank, world, lrank, make_halo, pkg.selfcheck.homogeneous_cases(world)) and ok
        ok = pkg.selfcheck.nrank_solver_and_actuators(pkg.capi, pkg.cases, rank, world, lrank, make_halo) and ok
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTIGPU_CHECK %s world=%d" % ("PASS" if t.item() == 1.0 else "FAIL", world), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
