"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: ncu_launches.py launches.csv"""
import csv
import sys
from collections import defaultdict
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith('==')) if r]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
t = defaultdict(float); n = defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    k = r[ki].split('(')[0]
    t[k] += v; n[k] += 1
tot = sum(t.values())
for k in sorted(t, key=t.get, reverse=True):
    print('%-28s n=%4d total=%10.1f us avg=%9.1f us share=%.3f' % (k, n[k], t[k] / 1e3, t[k] / n[k] / 1e3, t[k] / tot))
