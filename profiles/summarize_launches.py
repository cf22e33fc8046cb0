#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv > summary.txt"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
H = rows[hdr]; rows = rows[hdr + 1:]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r'^void\s+', '', r[ki]); name = re.sub(r'k_box<(.*)>\(T1, Box\)', r'\1', name)
    v = float(r[vi].replace(',', '')) * {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'nsecond': 1e-6, 'ms': 1.0, 'msecond': 1.0}.get(r[ui], 1e-6)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print('%d launches, %.3f ms total device time (cold-cache, serialised under ncu: compare SHARES)' % (len(rows), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print('%-44s n=%4d total %9.3f ms  avg %8.4f ms  share %5.1f%%' % (k[:44], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
