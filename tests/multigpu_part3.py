"""torchrun script: only the all-reduce paths of vfs-wind_b200/selfcheck.py (solver, Calc_U_lagr, Calc_F_eul, Pressure_Gradient)."""
import os
import sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_common as pc  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
pkg = pc.load_package()
ok = pkg.selfcheck.nrank_solver_and_actuators(pkg.capi, pkg.cases, rank, world, lrank, lambda c, cf: c.nccl_init(dist, device=torch.device("cuda", lrank)))
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_PART3 %s world=%d" % ("PASS" if t.item() == 1.0 else "FAIL", world), flush=True)
dist.destroy_process_group()
