#!/usr/bin/env python
"""Per-instruction stall samples of the hottest loop of a kernel (the instructions sharing the modal 'Instructions Executed'
count of the DMMAs), scaled to cycles per loop iteration.
usage: ncu_loop.py report.ncu-rep cycles_per_iteration [opcode=DMMA]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]; cyc = float(sys.argv[2]); opc = sys.argv[3] if len(sys.argv) > 3 else 'DMMA'
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ci, ce, cs = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall = [(i, c[6:]) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = [(i, r[ci].strip(), int(r[ce]), float(r[cs]), r) for i, r in enumerate(rows[2:]) if r[ce].isdigit()]
cnt = Counter(d[2] for d in data if opc in d[1])
modal = cnt.most_common(1)[0][0]
idx = [d[0] for d in data if d[2] == modal and opc in d[1]]
lo, hi = min(idx), max(idx)
loop = [d for d in data if lo - 2 <= d[0] <= hi + 3 and d[2] >= modal * 0.9]
tot = sum(d[3] for d in loop)
print('loop lines %d..%d, %d instructions, exec %d each, %d samples' % (lo, hi, len(loop), modal, tot))
byop = Counter(); bystall = Counter()
for d in loop:
    c = d[3] / tot * cyc
    top = sorted(((float(d[4][i]), n) for i, n in stall), reverse=True)[:3]
    for v, n in top: bystall[n] += v / tot * cyc
    op = d[1].split()[0]
    byop[op.split('.')[0]] += c
    print('%6d %-52s %6.1f cyc  %s' % (d[0], d[1][:52], c, ' '.join('%s:%.1f' % (n, v / tot * cyc) for v, n in top if v)))
print('by opcode (cycles/iteration):', ' '.join('%s %.1f' % kv for kv in byop.most_common()))
print('by stall  (cycles/iteration):', ' '.join('%s %.1f' % kv for kv in bystall.most_common()))
