#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, average, total, share."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, x in enumerate(rows) if x and x[0] == 'ID'][0]
hdr = rows[i]; agg = collections.defaultdict(list)
for x in rows[i + 1:]:
    if len(x) < len(hdr): continue
    name = x[hdr.index('Kernel Name')]
    m = re.search(r'(\w+)(?:<[^(]*)?\(', name); nm = m.group(1) if m else name[:40]
    agg[nm].append(float(x[hdr.index('Metric Value')]) / 1e6)
tot = sum(sum(v) for v in agg.values())
print("# %d launches captured, %.1f ms total" % (sum(len(v) for v in agg.values()), tot))
print("%-40s %6s %10s %10s %7s" % ("kernel", "n", "avg ms", "total ms", "share"))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-40s %6d %10.3f %10.2f %6.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
