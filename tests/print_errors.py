"""Print the per-field relative errors (GPU path vs oracle/_ref) of the parity cases: python tests/print_errors.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import parity_common as pc
import refdrv
pkg = pc.load_package()
worst = {}
for name, dims in (("c2_box256", (21, 17, 25)), ("c2_box256", (40, 33, 37)), ("c3_turbine", (29, 21, 25)), ("c3_turbine", (45, 30, 41)), ("c1_test10", (24, 16, 20)),
                   ("c1_test10", (60, 40, 50)), ("c2_box256", (101, 67, 50))):
    cfg = pkg.cases.scaled(pkg.cases.CONFIGS[name], *dims)
    err = pc.run_parity(cfg, refdrv, device=0)
    print(name, dims, {k: "%.1e" % v for k, v in err.items() if v > 1e-15}, flush=True)
    for k, v in err.items():
        worst[k] = max(worst.get(k, 0), v)
print("WORST", {k: "%.1e" % v for k, v in worst.items()})
