"""Aggregates an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
instructions executed (warp-level), thread instructions, stall samples.  usage: ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
cur_file = ''
for r in rows:
    if r and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) < 8:
        continue
    if r[0] != '' and r[2] == '-':          # a CUDA line header: totals for the line
        try:
            line = int(r[0])
        except ValueError:
            continue
        def g(name):
            try:
                return float(r[ci[name]])
            except Exception:
                return 0.0
        a = agg.setdefault((cur_file, line), [r[1].strip()[:90], 0, 0, 0])
        a[1] += g('Instructions Executed'); a[2] += g('Thread Instructions Executed'); a[3] += g('# Samples')
tot_i = sum(a[1] for a in agg.values()); tot_s = sum(a[3] for a in agg.values()); tot_t = sum(a[2] for a in agg.values())
print('total warp-inst %.3g thread-inst %.3g samples %d  avg active %.2f' % (tot_i, tot_t, tot_s, tot_t / max(tot_i, 1)))
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%-22s %4d  inst %5.2f%%  samp %5.2f%%  act %5.1f  | %s' % (line[0][:22], line[1], 100 * a[1] / tot_i, 100 * a[3] / max(tot_s, 1), a[2] / max(a[1], 1), a[0]))

if len(sys.argv) > 3:   # region summary: file:lo-hi=name,...
    for spec in sys.argv[3].split(','):
        rng, name = spec.split('=')
        f, lh = rng.split(':')
        lo, hi = map(int, lh.split('-'))
        sel = [a for (ff, ll), a in agg.items() if ff.startswith(f) and lo <= ll <= hi]
        print('%-28s inst %5.2f%%  samp %5.2f%%  act %5.1f' % (name, 100 * sum(a[1] for a in sel) / tot_i, 100 * sum(a[3] for a in sel) / max(tot_s, 1),
                                                              sum(a[2] for a in sel) / max(sum(a[1] for a in sel), 1)))
