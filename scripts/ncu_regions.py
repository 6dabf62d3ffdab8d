#!/usr/bin/env python
"""Split the stall samples of a warp-specialised kernel into producer / consumer regions.
usage: ncu_regions.py report.ncu-rep"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); m = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum']:
    print('%-85s %s' % (k, m.get(k)))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ci, ce, cs = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
stall = [(i, c[6:]) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = [(i, r[ci].strip(), int(r[ce]), float(r[cs]), r) for i, r in enumerate(rows[2:]) if r[ce].isdigit()]
def op(t):
    p = t.split(); return (p[1] if p[0].startswith('@') else p[0]).split('.')[0]
split = [d[0] for d in data if 'USETMAXREG.TRY_ALLOC' in d[1]]
split = split[0] if split else len(data)
for name, seg in (('producer', [d for d in data if d[0] < split]), ('consumer', [d for d in data if d[0] >= split])):
    tot = sum(d[3] for d in seg) or 1
    c = Counter(); o = Counter(); ex = Counter()
    for d in seg:
        for i, n in stall: c[n] += float(d[4][i])
        o[op(d[1])] += d[3]; ex[op(d[1])] += d[2]
    print('== %s: samples %d, warp-instr %.3e' % (name, tot, sum(d[2] for d in seg)))
    print('   stalls  :', ' '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in c.most_common(7)))
    print('   samples :', ' '.join('%s %.1f%%' % (k, 100 * v / tot) for k, v in o.most_common(8)))
    print('   executed:', ' '.join('%s %.2e' % kv for kv in ex.most_common(10)))
    for d in sorted(seg, key=lambda t: -t[3])[:int(sys.argv[2]) if len(sys.argv) > 2 else 8]:
        top = sorted(((float(d[4][i]), n) for i, n in stall), reverse=True)[:2]
        print('   %5.2f%% #%d %-46s exec %-10d %s' % (100 * d[3] / tot, d[0], d[1][:46], d[2], ' '.join('%s:%d' % (n, v) for v, n in top if v)))
