#!/usr/bin/env python
"""ncu raw page CSV (ncu -i X.ncu-rep --page raw --csv) -> one short text summary per kernel under profiles/.
usage: python tools/ncu_summary.py <raw.csv> <prefix, e.g. profiles/r02_> <what was run>"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct']
stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
for r in rows[2:]:
    m = re.search(r'(\w+)(?:<[^(]*)?\(', r[idx['Kernel Name']])
    short = m.group(1) if m else r[idx['Kernel Name']][:30]
    with open('%s%s_ncu.txt' % (sys.argv[2], short), 'w') as o:
        o.write('# ncu --set full --clock-control none, one launch of %s inside `%s` (B200), round 2\n' % (short, sys.argv[3]))
        o.write('# (cold-cache, serialised replay: durations are for shares, the bench line carries the live CUDA-event times)\n')
        for k in keys:
            if k in idx:
                o.write('%-78s %s %s\n' % (k, r[idx[k]], units[idx[k]]))
        o.write('# warps stalled per issue-active cycle, top reasons\n')
        for v, h in sorted([(float(r[idx[h]] or 0), h) for h in stall], reverse=True)[:6]:
            o.write('   %6.2f %s\n' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print(short)
