#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line and by
function region.  usage: ncu_by_line.py dump.csv file_substr [region_start_line:name ...]"""
import csv, sys, bisect
def num(x):
    try: return int(float(x))
    except ValueError: return 0
path, fsub = sys.argv[1], sys.argv[2]
regions = sorted((int(a.split(':')[0]), a.split(':')[1]) for a in sys.argv[3:])
rows = list(csv.reader(open(path)))
cur = None; hdr = None; lines = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur = r[1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; iI = r.index('Instructions Executed'); iS = r.index('# Samples'); iT = r.index('Thread Instructions Executed'); continue
    if hdr is None or cur is None: continue
    if r[0] != '':
        try: ln = int(r[0])
        except ValueError: continue
        key = (cur, ln)
        if key not in lines: lines[key] = [r[1], 0, 0, 0, 0]
        continue
    if r[2].startswith('0x'):   # SASS rows follow their CUDA line
        lines[key][1] += num(r[iI]); lines[key][2] += num(r[iS]); lines[key][3] += 1; lines[key][4] += num(r[iT])
tot = sum(v[1] for v in lines.values()); tots = sum(v[2] for v in lines.values()); nsass = sum(v[3] for v in lines.values())
print('total warp inst %d  samples %d  sass instrs %d' % (tot, tots, nsass))
byfile = {}
for (f, ln), v in lines.items():
    a = byfile.setdefault(f, [0, 0, 0]); a[0] += v[1]; a[1] += v[2]; a[2] += v[3]
for f, a in byfile.items(): print('  %-60s inst %5.1f%% samples %5.1f%% sass %d' % (f[-60:], 100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2]))
if regions:
    agg = {}
    starts = [x[0] for x in regions]
    for (f, ln), v in lines.items():
        if fsub not in f: continue
        k = bisect.bisect_right(starts, ln) - 1
        name = regions[k][1] if k >= 0 else 'pre'
        a = agg.setdefault(name, [0, 0, 0, 0]); a[0] += v[1]; a[1] += v[2]; a[2] += v[3]; a[3] += v[4]
    for k, v in agg.items(): print('%-14s inst %5.1f%%  samples %5.1f%%  sass %5d  lanes/inst %.1f' % (k, 100 * v[0] / tot, 100 * v[1] / max(tots, 1), v[2], v[3] / max(v[0], 1)))
print('--- top lines by samples')
for (f, ln), v in sorted(lines.items(), key=lambda x: -x[1][2])[:40]:
    print('%s:%4d inst %5.2f%% smp %5.2f%% sass %4d | %s' % (f.split('/')[-1][:16], ln, 100 * v[1] / tot, 100 * v[2] / max(tots, 1), v[3], v[0].strip()[:100]))
