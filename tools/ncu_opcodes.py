"""Dynamic (executed) warp-level instruction histogram by SASS opcode from an `ncu --page source --csv --print-source cuda,sass` dump."""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
cnt = Counter()
for r in rows:
    if r and r[0] == 'Line No':
        hdr = r
        ci = {n: i for i, n in enumerate(hdr)}
        continue
    if hdr is None or len(r) < 8 or r[0] != '' or r[2] in ('-', '...'):
        continue
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[3])
    if not m:
        continue
    try:
        n = float(r[ci['Instructions Executed']])
    except ValueError:
        continue
    cnt[m.group(2)] += n
tot = sum(cnt.values())
print('total %.4g' % tot)
for op, n in cnt.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30):
    print('%-10s %6.2f%%  %.4g' % (op, 100 * n / tot, n))
