"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck), e.g.
    compute-sanitizer --tool racecheck python tests/sanitize_small.py
Exercises every marching kernel of the default path on several tiles and k-chunks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import parity_common as pc
pkg = pc.load_package()
capi, cases = pkg.capi, pkg.cases
for name, dims in (("c2_box256", (70, 37, 40)), ("c3_turbine", (45, 30, 35))):
    cfg = cases.scaled(cases.CONFIGS[name], *dims)
    mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
    ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
    ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
    met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
    f = cases.make_fields(cfg, met)
    for k, n in pc.FIELDS_IN:
        ctx.upload(n, f[k])
    ctx.rhs_les_fused()
    ctx.Convection(); ctx.Viscous()
    print(name, float(np.abs(ctx.download("RHS")).max()), flush=True)
    ctx.close()
