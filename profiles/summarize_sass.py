"""Regenerate profiles/sass_summary.txt: SASS mnemonic counts per kernel of the product library.
Usage: python profiles/summarize_sass.py  (needs cuobjdump + c++filt; no GPU)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "vfs-wind_b200", "libvfs_b200.so")
COLS = [("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("SHFL", r"\bSHFL"), ("DFMA", r"\bDFMA"), ("DMUL", r"\bDMUL"), ("DADD", r"\bDADD"),
        ("LDG.128", r"\bLDG\.E\.128|\bLDG\.E\.[A-Z.]*128"), ("LDG", r"\bLDG"), ("LDS", r"\bLDS"), ("BAR", r"\bBAR\.")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            continue
        if name is None or "/*" not in line:
            continue
        for col, rx in COLS:
            if re.search(rx, line):
                counts[name][col] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    rows = sorted(zip(dem, names), key=lambda r: -counts[r[1]]["DFMA"])
    out = ["# SASS mnemonic counts per kernel of vfs-wind_b200/libvfs_b200.so (cuobjdump -sass), sm_100a, nvcc 12.9; regenerate: python profiles/summarize_sass.py",
           "# UTMALDG = cp.async.bulk.tensor (TMA loads), SYNCS = mbarrier ops, SHFL = warp shuffles, DFMA/DMUL/DADD = FP64 pipe, LDG.128 = 16-byte global loads",
           "# no HMMA / UTCMMA / tcgen05 anywhere: the path is an FP64 stencil, not a contraction",
           "%-90s" % "kernel" + "".join("%8s" % c for c, _ in COLS)]
    tc = 0
    for d, n in rows:
        out.append("%-90s" % d[:90] + "".join("%8d" % counts[n][c] for c, _ in COLS))
    tc = len(re.findall(r"\b(HMMA|UTCMMA|UTCHMMA|IMMA|DMMA)\b", sass))
    out.append("# tensor-core instructions in the whole library: %d" % tc)
    open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
    print("wrote sass_summary.txt: %d kernels" % len(rows))


if __name__ == "__main__":
    main()
