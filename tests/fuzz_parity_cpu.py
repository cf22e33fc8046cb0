"""Randomised parity sweep on the CPU: kernel logic (host emulation, test-only) against the oracle over random
boundary-type / flag combinations on small grids.  python tests/fuzz_parity_cpu.py [n] [seed]
FUZZ_GPU=1: the same sweep through the CUDA library on cuda:0 instead of the emulation (FUZZ_BIG=1: multi-tile sizes)."""
import os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import parity_common as pc, emu_loader, refdrv
pkg = pc.load_package(); emu = None if os.environ.get("FUZZ_GPU") else emu_loader.load(pkg.capi)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
WALL_I = [1, 10, -1, -2, 100]
WALL_JLO = [1, 10, 12, -1, -2]
WALL_JHI = [1, 2, 4, 10, -10, 12, 13, 14, -1, -2]
bad_cases = []
for t in range(n):
    if os.environ.get("FUZZ_BIG"):     # several marching tiles wide and high (tile strides 30 x 10/14), several k-chunks
        dims = (rng.randint(31, 45), rng.randint(13, 20), rng.randint(17, 22))
    else:
        dims = (rng.randint(9, 15), rng.randint(8, 13), rng.randint(9, 14))
    base = pkg.cases.scaled(pkg.cases.CONFIGS[rng.choice(["c3_turbine", "c2_box256"])], *dims)
    per = [rng.random() < 0.4, rng.random() < 0.15, rng.random() < 0.4]
    bc = [100] * 6
    if not per[0]: bc[0], bc[1] = rng.choice(WALL_I), rng.choice(WALL_I)
    if not per[1]: bc[2], bc[3] = rng.choice(WALL_JLO), rng.choice(WALL_JHI)
    if not per[2]: bc[4], bc[5] = rng.choice([1, 5, 100]), rng.choice([1, 4, 100])
    leg = [rng.random() < 0.3 for _ in range(3)]      # a periodic direction through the legacy i/j/k_periodic switch instead of the DA wrap
    fl = dict(base["flags"], ii_periodic=int(per[0] and not leg[0]), jj_periodic=int(per[1] and not leg[1]), kk_periodic=int(per[2] and not leg[2]),
              i_periodic=int(per[0] and leg[0]), j_periodic=int(per[1] and leg[1]), k_periodic=int(per[2] and leg[2]),
              second_order=rng.randint(0, 1), laplacian=rng.randint(0, 1), immersed=rng.choice([0, 1, 3]),
              les=rng.choice([0, 1, 2, 2]), testfilter_ik=int(rng.random() < 0.15), roughness_size=1e-3,
              rotor_model=rng.randint(0, 1))
    # the variants of the one-thread-per-face / per-cell kernels (momentum.c:754-923, les.c:420-656, 798-965, 1211)
    if rng.random() < 0.25: fl["skew"] = 1
    if rng.random() < 0.25: fl["clark"] = 1
    r = rng.random()
    if r < 0.12: fl["inviscid"] = 1
    elif r < 0.3: fl["levelset_weno"] = 5
    if rng.random() < 0.15: fl["wallfunction"] = 2
    if fl["les"] == 2 and rng.random() < 0.2: fl.update(rng.choice([dict(i_homo_filter=1, k_homo_filter=1), dict(i_homo_filter=1), dict(j_homo_filter=1), dict(k_homo_filter=1)]))
    if not per[0] and rng.random() < 0.15: bc[0] = 11      # body-fitted cylinder inflow half (rhs.c:626): needs z <= 0 somewhere
    if rng.random() < 0.2: fl.update(ti=5, tistart=5)
    if bc[2] in (1,) and rng.random() < 0.3: fl["viscosity_wallmodel"] = 1
    cfg = dict(base); cfg["flags"] = fl; cfg["bctype"] = bc
    if bc[0] == 11: cfg["z_shift"] = 1.4
    try:
        err = pc.run_parity(cfg, refdrv, lib=emu)
    except Exception as e:      # unsupported combination rejected by vfs_create, etc.
        print(t, "SKIP", bc, {k: v for k, v in fl.items() if v}, str(e)[:80]); continue
    zp = err.pop("FormFunction_SNES_zero_pattern"); nvm = err.pop("IB_BC_nvert_mismatches")
    bad = {k: v for k, v in err.items() if not (v <= 1e-12)}
    if bad or zp or nvm:
        bad_cases.append((bc, fl, bad, zp, nvm))
        print(t, "FAIL", bc, {k: v for k, v in fl.items() if v}, (cfg["IM"], cfg["JM"], cfg["KM"]), {k: "%.1e" % v for k, v in bad.items()}, zp, nvm, flush=True)
    else:
        print(t, "ok", bc, flush=True)
print("failures:", len(bad_cases))
