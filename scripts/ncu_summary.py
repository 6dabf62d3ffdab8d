#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics of the first kernel + hottest SASS lines by stall samples.
usage: ncu_summary.py report.ncu-rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = dict(zip(hdr, zip(units, vals)))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in m:
        print('%-80s %s %s' % (k, m[k][1], m[k][0]))
print('--- stall reasons (warps per issue-active cycle)')
for k in sorted(hdr):
    if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
        v = float(m[k][1])
        if v > 0.05:
            print('  %-40s %.3f' % (k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci, cs, ce = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = []
for idx, r in enumerate(rows[2:]):
    try:
        data.append((float(r[cs]), idx, r[ci].strip(), int(r[ce]), r))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
print('--- total samples %d over %d SASS lines; hottest lines' % (tot, len(data)))
for s_, idx, text, ex, r in sorted(data, key=lambda t: -t[0])[:nl]:
    top = sorted(((float(r[i]), c[6:]) for i, c in stall_cols), reverse=True)[:2]
    print('%6.2f%%  #%5d  %-62s exec %-10d %s' % (100 * s_ / tot, idx, text[:62], ex, ' '.join('%s:%d' % (c, v) for v, c in top if v)))
# coarse regions: cumulative samples by 5%% of the instruction stream
print('--- samples by instruction-index decile')
n = len(data)
for d in range(10):
    seg = [x for x in data if d * n // 10 <= x[1] < (d + 1) * n // 10]
    print('  lines %5d-%5d: %5.1f%%' % (d * n // 10, (d + 1) * n // 10, 100 * sum(x[0] for x in seg) / tot))
