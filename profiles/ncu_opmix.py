#!/usr/bin/env python
"""Dynamic opcode mix of a kernel from an `ncu --page source --print-source cuda,sass --csv` dump.  usage: ncu_opmix.py dump.csv"""
import csv, re, sys, collections
agg = collections.Counter(); smp = collections.Counter(); tot = 0; tots = 0
for r in csv.reader(open(sys.argv[1])):
    if len(r) < 8 or r[0] != '' or not r[2].startswith('0x'):
        continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)', r[3])
    if not m:
        continue
    try:
        n = int(float(r[7])); s = int(float(r[6]))
    except ValueError:
        continue
    op = m.group(2); base = op.split('.')[0]
    if base == 'IMAD' and ('.MOV' in op): base = 'IMAD.MOV'
    agg[base] += n; smp[base] += s; tot += n; tots += s
print('warp instructions %d, samples %d' % (tot, tots))
for k, v in agg.most_common(45):
    print('%-12s inst %6.2f%%  samples %6.2f%%' % (k, 100 * v / tot, 100 * smp[k] / max(tots, 1)))
