#!/usr/bin/env python
"""Aggregate an .ncu-rep source page by CUDA source line: instructions executed and stall samples (top N lines)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    key = (cur, int(r[0]))
    try:
        inst = int(r[hdr.index("Instructions Executed")] or 0); smp = int(r[hdr.index("# Samples")] or 0); tinst = int(r[hdr.index("Thread Instructions Executed")] or 0)
    except ValueError:
        continue
    a = agg.setdefault(key, [0, 0, 0, r[1]])
    a[0] += inst; a[1] += smp; a[2] += tinst
ti = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp-inst %d, samples %d" % (ti, ts))
print("| file:line | inst % | samples % | avg thr | source |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print("| %s:%d | %.1f | %.1f | %.1f | `%s` |" % (k[0], k[1], 100.0 * a[0] / ti, 100.0 * a[1] / ts, a[2] / max(1, a[0]), a[3].strip()[:110]))
