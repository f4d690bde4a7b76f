#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    if hdr is None:
        if 'Kernel Name' in r:
            hdr = r
        continue
    d = dict(zip(hdr, r))
    if d.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    n = re.sub(r'\(.*', '', d['Kernel Name'])
    n = re.sub(r'cub::CUB_\w+::', 'cub::', n)
    v = float(d['Metric Value'].replace(',', ''))
    u = d['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v)
    a = agg[n]
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | max ms | share |\n|---|---|---|---|---|")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("| `%s` | %d | %.2f | %.2f | %.1f%% |" % (n[:80], a[0], a[1] / 1e3, a[2] / 1e3, 100 * a[1] / tot))
print("| total | | %.2f | | |" % (tot / 1e3))
