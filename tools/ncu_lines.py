#!/usr/bin/env python
"""Per source line: instructions executed and stall samples of one kernel, from the correlated source page of an ncu report taken with
--import-source on:  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > page.csv
usage: python tools/ncu_lines.py <page.csv> [top N] [--buckets]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
cur, hdr, per = None, None, collections.defaultdict(lambda: [0, 0, ""])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":          # a source line (its SASS rows follow, with addresses)
        try:
            e = per[(cur, int(r[0]))]
            e[0] += int(r[hdr.index("Instructions Executed")]); e[1] += int(r[hdr.index("# Samples")]); e[2] = r[1].strip()[:110]
        except ValueError:
            pass
ti = sum(e[0] for e in per.values()); ts = sum(e[1] for e in per.values())
print("warp instructions %d, samples %d" % (ti, ts))
for (f, ln), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100.0 * e[0] / ti, 100.0 * e[1] / max(ts, 1), f, ln, e[2]))
if "--buckets" in sys.argv:
    bk = collections.defaultdict(lambda: [0, 0])
    for (f, ln), e in per.items():
        b = (f, ln // 25 * 25); bk[b][0] += e[0]; bk[b][1] += e[1]
    print("--- 25-line buckets")
    for (f, b), e in sorted(bk.items()):
        if e[0] * 200 > ti:
            print("%5.2f%% inst %5.2f%% smp  %s:%d-%d" % (100.0 * e[0] / ti, 100.0 * e[1] / max(ts, 1), f, b, b + 24))
