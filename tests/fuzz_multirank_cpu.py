"""Randomised N-rank == 1-rank sweep on the CPU (gloo, kernel logic through the test-only host emulation): random
boundary types / flags / slab counts; every downloaded field must be BITWISE equal to the single-rank run.
python tests/fuzz_multirank_cpu.py [n] [seed]"""
import os, sys, random, socket, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch.multiprocessing as mp
import parity_common as pc, emu_loader
import test_cpu_multirank as tm


def main():
    pkg = pc.load_package(); lib = emu_loader.load(pkg.capi)
    capi, cases = pkg.capi, pkg.cases
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    fails = 0
    for t in range(n):
        world = rng.choice([2, 2, 3])
        name = rng.choice(["c3_turbine", "c2_box256"])
        dims = (rng.randint(9, 14), rng.randint(8, 12), rng.randint(4 * world + 3, 4 * world + 9))
        per = [rng.random() < 0.4, rng.random() < 0.15, rng.random() < 0.5]
        bc = [100] * 6
        if not per[0]: bc[0], bc[1] = rng.choice([1, 10, -1, -2]), rng.choice([1, 10, -1, -2])
        if not per[1]: bc[2], bc[3] = rng.choice([1, 10, 12, -1, -2]), rng.choice([1, 2, 4, 10, -10, 12, 13, 14, -1, -2])
        if not per[2]: bc[4], bc[5] = rng.choice([1, 5]), rng.choice([1, 4])
        leg = [rng.random() < 0.3, rng.random() < 0.3]      # i / j periodic through the legacy switches (k_periodic is single-rank only)
        extra = dict(ii_periodic=int(per[0] and not leg[0]), jj_periodic=int(per[1] and not leg[1]), kk_periodic=int(per[2]),
                     i_periodic=int(per[0] and leg[0]), j_periodic=int(per[1] and leg[1]), second_order=rng.randint(0, 1),
                     laplacian=rng.randint(0, 1), immersed=rng.choice([0, 1, 3]), les=rng.choice([0, 1, 2, 2]), roughness_size=1e-3)
        if rng.random() < 0.3: extra["skew"] = 1          # Adv1-3 and the Clark gradient planes live outside the main pool and travel too
        if rng.random() < 0.3: extra["clark"] = 1
        r = rng.random()
        if r < 0.12: extra["inviscid"] = 1
        elif r < 0.3: extra["levelset_weno"] = 5
        if rng.random() < 0.25: extra.update(ti=5, tistart=5)
        cfg = cases.scaled(cases.CONFIGS[name], *dims)
        cfg["flags"] = dict(cfg["flags"], **extra); cfg["bctype"] = bc
        mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
        with tempfile.TemporaryDirectory() as tmp:
            xyz = cases.make_grid(cfg)
            ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]), lib=lib)
            ctx.upload("COOR", xyz); ctx.FormMetrics()
            met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
            f = cases.make_fields(cfg, met)
            for k, nm in pc.FIELDS_IN: ctx.upload(nm, f[k])
            x = f["ucont"] * (1.0 + 1e-3 * np.sin(np.arange(f["ucont"].size).reshape(f["ucont"].shape)))
            single = pc.run_path(ctx, x); single["NVERT"] = ctx.download("NVERT"); ctx.close()
            np.savez(os.path.join(tmp, "global.npz"), xyz=xyz, x=x, **f)
            s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
            mp.spawn(tm._worker, args=(world, port, tmp, name, dims, extra, bc), nprocs=world, join=True)
            parts = [np.load(os.path.join(tmp, "rank%d.npz" % r)) for r in range(world)]
        bad = [nm for nm in ("F", "UCAT", "CS", "NU_T", "UCONT", "FUSED_RHS", "FUSED_UCAT", "FUSED_CS", "FUSED_NU_T", "PROJ_P", "PROJ_PHI", "PROJ_UCONT")
               if not np.array_equal(np.concatenate([pp[nm] for pp in parts], axis=0), single[nm], equal_nan=True)]
        print(t, "FAIL" if bad else "ok", world, name, dims, bc, {k: v for k, v in extra.items() if v}, bad, flush=True)
        fails += bool(bad)
    print("failures:", fails)


if __name__ == "__main__":
    main()
