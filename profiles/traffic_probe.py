"""Workload for the per-launch DRAM traffic capture behind profiles/traffic.json:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python profiles/traffic_probe.py
    python profiles/make_traffic_json.py gpurun_out/traffic.csv <launches per step printed by the probe>
Runs the RHS+LES unit of config 2 (256^3) three times eagerly (no CUDA graph); the LAST step's launches are the sample."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package
pkg = load_package(); capi, cases = pkg.capi, pkg.cases
cfg = dict(cases.CONFIGS["c2_box256"])
mx, my, mz = cfg["IM"] + 1, cfg["JM"] + 1, cfg["KM"] + 1
ctx = capi.VfsContext(capi.make_params(mx, my, mz, cfg["flags"], cfg["ren"], cfg["dt"], cfg["bctype"]))
ctx.upload("COOR", cases.make_grid(cfg)); ctx.FormMetrics()
met = dict(csi=ctx.download("CSI"), eta=ctx.download("ETA"), zet=ctx.download("ZET"), aj=ctx.download("AJ"))
f = cases.make_fields(cfg, met)
for k, n in (("nvert", "NVERT"), ("ucont", "UCONT"), ("ucat", "UCAT"), ("ucat_old", "UCAT_OLD"), ("ucont_o", "UCONT_O"), ("rhs_o", "RHS_O"), ("dp", "DP"), ("f_eul", "F_EUL")):
    ctx.upload(n, f[k])
ctx.set_option(1, 0)
for it in range(3):
    l0 = ctx.launch_count()
    ctx.rhs_les_fused()
    print("launches in step %d: %d" % (it, ctx.launch_count() - l0), flush=True)
ctx.close()
