#!/usr/bin/env python
"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch of ONE step) -> profiles/r02_spike_traffic.json:
DRAM bytes per step over every kernel, and per kernel name.
usage: python tools/ncu_traffic.py <csv> <out.json> [--one-step] [--last-id N] [--add name read_bytes write_bytes ms]...
--one-step keeps the launches from the first parse_kernel of the capture up to (not including) the next one = one pass over the body;
--last-id drops launches after ID N (the capture window ran on into the next phase of the bench); --add supplies a launch the window missed
from another capture of the same command (the --set full report of that kernel)."""
import collections, csv, json, re, sys
rows = list(csv.reader(open(sys.argv[1])))
last_id = int(sys.argv[sys.argv.index("--last-id") + 1]) if "--last-id" in sys.argv else None
adds = [sys.argv[k + 1:k + 5] for k, a in enumerate(sys.argv) if a == "--add"]
i = [k for k, x in enumerate(rows) if x and x[0] == "ID"][0]
hdr = rows[i]
first_id = None
if "--one-step" in sys.argv:
    ids = sorted({int(x[0]) for x in rows[i + 1:] if len(x) >= len(hdr) and "parse_kernel" in x[hdr.index("Kernel Name")]})
    assert len(ids) >= 2, "the capture must hold two parse_kernel launches"
    first_id, last_id = ids[0], ids[1] - 1
per = collections.defaultdict(lambda: {"launches": set(), "read": 0.0, "write": 0.0, "ms": 0.0})
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}
for x in rows[i + 1:]:
    if len(x) < len(hdr):
        continue
    if last_id is not None and int(x[0]) > last_id:
        continue
    if first_id is not None and int(x[0]) < first_id:
        continue
    name = x[hdr.index("Kernel Name")]
    m = re.search(r"(\w+)(?:<[^(]*)?\(", name)
    nm = m.group(1) if m else name[:40]
    metric, unit, val = x[hdr.index("Metric Name")], x[hdr.index("Metric Unit")], float(x[hdr.index("Metric Value")].replace(",", ""))
    e = per[nm]
    e["launches"].add(x[0])
    if metric == "dram__bytes_read.sum":
        e["read"] += val * scale.get(unit, 1.0)
    elif metric == "dram__bytes_write.sum":
        e["write"] += val * scale.get(unit, 1.0)
    elif metric == "gpu__time_duration.sum":
        e["ms"] += val * scale.get(unit, 1.0)
for nm, rd, wr, ms in adds:
    e = per[nm]
    e["launches"].add("added")
    e["read"] += float(rd); e["write"] += float(wr); e["ms"] += float(ms)
out = {"notes": (["launches %s..%d of the capture = one step" % (first_id, last_id)] if first_id is not None else ["launches after ID %d dropped (next phase of the bench)" % last_id] if last_id is not None else []) +
                ["%s taken from its own --set full capture of the same command" % a[0] for a in adds],
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over the launches of one step of `python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-tnc` (C2)",
       "dram_bytes_per_step": sum(e["read"] + e["write"] for e in per.values()),
       "per_kernel": {k: {"launches": len(e["launches"]), "dram_read_bytes": e["read"], "dram_write_bytes": e["write"], "ms": round(e["ms"], 4)}
                      for k, e in sorted(per.items(), key=lambda kv: -(kv[1]["read"] + kv[1]["write"]))}}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("dram bytes per step: %.3f GB over %d kernels" % (out["dram_bytes_per_step"] / 1e9, len(per)))
