"""profiles/traffic.json from the ncu CSV of profiles/traffic_probe.py: DRAM bytes and time of the LAST `n` launches (one
RHS+LES step), per kernel group and summed.  usage: make_traffic_json.py traffic.csv n [label]"""
import collections, csv, json, os, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
H = rows[[i for i, r in enumerate(rows) if r[0] == "ID"][0]]
rows = [r for r in rows if r[0] != "ID"]
n = int(sys.argv[2])
ki, mi, vi, ui, ii = H.index("Kernel Name"), H.index("Metric Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(int(r[ii]), {"name": r[ki], "bytes": 0.0, "ms": 0.0})
    v = float(r[vi].replace(",", ""))
    if r[mi].startswith("dram__bytes"):
        d["bytes"] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
    elif r[mi].startswith("gpu__time_duration"):
        d["ms"] += v * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}.get(r[ui], 1e-6)
ids = sorted(launch)[-n:]
GROUPS = (("c2c", "C2CInterior"), ("les1", "Les1Body"), ("les2", "k_les2_march"), ("les3", "Les3March"), ("nut", "NuT<"), ("flux", "k_flux_march"),
          ("fp", "FpCell"), ("project", "ProjectSNES"))
OTHER = "other (ghost refreshes, boundary shells, thin slabs)"
by_b, by_ms, cnt, per_launch = collections.OrderedDict(), collections.OrderedDict(), collections.OrderedDict(), {}
for i in ids:
    d = launch[i]
    g = next((g for g, pat in GROUPS if pat in d["name"]), OTHER)
    if g in ("nut", "les3", "fp", "project", "c2c") and d["bytes"] < 2e8:
        g = OTHER                         # thin-slab launches of the same functor
    by_b[g] = by_b.get(g, 0.0) + d["bytes"]; by_ms[g] = by_ms.get(g, 0.0) + d["ms"]; cnt[g] = cnt.get(g, 0) + 1
for g in by_b:
    if g != OTHER:
        per_launch[g] = by_b[g] / cnt[g]
total = sum(by_b.values())
alg = 248.0 * 254 ** 3
out = {"workload": "c2_box256",
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over profiles/traffic_probe.py: "
              "the %d launches of ONE vfs_rhs_les_fused step (eager, no graph), B200%s" % (n, (", " + sys.argv[3]) if len(sys.argv) > 3 else ""),
       "dram_bytes_per_launch": per_launch, "launches_per_step": cnt, "dram_bytes_per_step_by_group": by_b,
       "ms_under_ncu_by_group": {g: round(v, 4) for g, v in by_ms.items()},
       "dram_bytes_per_step": total, "algorithmic_bytes_per_step": alg, "wasted_traffic_ratio": total / alg}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1)
print(json.dumps({"dram_bytes_per_step": total, "ratio": total / alg, "launches": cnt}, indent=1))
