#!/usr/bin/env python
"""Stall reasons by function region from an `ncu --page source --print-source cuda,sass --csv` dump.
usage: ncu_stalls.py dump.csv file_substr start:name ..."""
import csv, sys, bisect
def num(x):
    try: return int(float(x))
    except ValueError: return 0
path, fsub = sys.argv[1], sys.argv[2]
regions = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[3:])
starts = [x[0] for x in regions]
rows = list(csv.reader(open(path)))
cur = None; hdr = None; key = None
agg = {}
cols = ['stall_wait', 'stall_short_sb', 'stall_long_sb', 'stall_barrier', 'stall_math', 'stall_no_inst', 'stall_branch_resolving', 'stall_not_selected', 'stall_selected', 'stall_dispatch', 'stall_mio', 'stall_lg']
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No':
        hdr = r; idx = [r.index(c) for c in cols]; iI = r.index('Instructions Executed'); continue
    if hdr is None or cur is None: continue
    if r[0] != '':
        try: key = (cur, int(r[0]))
        except ValueError: key = None
        continue
    if key and r[2].startswith('0x'):
        if fsub in key[0]:
            k = bisect.bisect_right(starts, key[1]) - 1
            name = regions[k][1] if k >= 0 else 'pre'
        else:
            name = 'other:' + key[0].split('/')[-1][:20]
        a = agg.setdefault(name, [0] * (len(cols) + 1))
        a[0] += num(r[iI])
        for j, i in enumerate(idx): a[j + 1] += num(r[i])
tot = [sum(a[j] for a in agg.values()) for j in range(len(cols) + 1)]
print('%-16s %7s ' % ('region', 'inst%') + ' '.join('%9s' % c[6:15] for c in cols))
for name, a in sorted(agg.items(), key=lambda x: -sum(x[1][1:])):
    s = sum(tot[1:])
    print('%-16s %6.1f%% ' % (name, 100 * a[0] / tot[0]) + ' '.join('%8.2f%%' % (100 * v / s) for v in a[1:]))
print('%-16s %6.1f%% ' % ('total', 100) + ' '.join('%8.2f%%' % (100 * v / sum(tot[1:])) for v in tot[1:]))
