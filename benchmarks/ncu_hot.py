#!/usr/bin/env python
"""Top stall-sample SASS instructions of an `ncu --page source --csv` export.  python benchmarks/ncu_hot.py file.csv [top] [context]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
body = [r for r in rows[2:] if len(r) >= 8]
tot = sum(int(r[2] or 0) for r in body)
texec = sum(int(r[5] or 0) for r in body)
print("instructions:", len(body), "samples:", tot, "warp-inst executed:", texec)
idx = sorted(range(len(body)), key=lambda i: -int(body[i][2] or 0))[:top]
for i in sorted(idx):
  r = body[i]
  print("%5d %6.2f%% smp=%6s exec=%10s thr=%5s  %s" % (i, 100.0 * int(r[2] or 0) / max(tot, 1), r[2], r[5], r[8], r[1].strip()[:100]))
