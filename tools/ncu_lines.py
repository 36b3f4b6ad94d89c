#!/usr/bin/env python
"""Join an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --kernel-name regex:K) with `nvdisasm -g -c` line info of the same
cubin, by instruction order: per source line instructions executed and stall samples.
usage: python tools/ncu_lines.py <page.csv> <nvdisasm.sass> <kernel substring> [top N]"""
import csv, re, sys, collections
page, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(page)))
h = [k for k, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ins = [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]
lines = open(sass).read().split("\n")
# the function's text section
start = [k for k, l in enumerate(lines) if l.startswith(".text.") and kern in l]
assert start, "kernel not in sass"
k = start[0] + 1
cur = ("?", 0)
seq = []
stack = ""
while k < len(lines) and not lines[k].startswith(".text.") and not lines[k].startswith("//--------------------- .") :
    l = lines[k]
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        stack = m.group(3)
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        seq.append((cur, stack, l.split("*/", 1)[1].strip()[:60]))
    k += 1
print("sass instructions", len(seq), "ncu rows", len(ins))
n = min(len(seq), len(ins))
per = collections.defaultdict(lambda: [0, 0, 0])
ti = ts = 0
for a, b in zip(seq[:n], ins[:n]):
    e = per[a[0]]
    i, s = int(b["Instructions Executed"]), int(b["# Samples"])
    e[0] += i; e[1] += s; e[2] += 1
    ti += i; ts += s
print("warp instructions %d, samples %d" % (ti, ts))
src = {}
for (f, ln), e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open("/root/repo/stochasticsim_b200/csrc/" + f).read().split("\n")
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ""
    print("%5.2f%% inst %5.2f%% smp %4d sass  %s:%d  %s" % (100.0 * e[0] / ti, 100.0 * e[1] / max(ts, 1), e[2], f, ln, text))
if "--buckets" in sys.argv:
    bk = collections.defaultdict(lambda: [0, 0])
    for (f, ln), e in per.items():
        b = (f, ln // 25 * 25)
        bk[b][0] += e[0]; bk[b][1] += e[1]
    print("--- 25-line buckets")
    for (f, b), e in sorted(bk.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if e[0] * 200 > ti:
            print("%5.2f%% inst %5.2f%% smp  %s:%d-%d" % (100.0 * e[0] / ti, 100.0 * e[1] / max(ts, 1), f, b, b + 24))
