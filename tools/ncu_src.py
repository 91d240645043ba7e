#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: stall-reason totals, samples per
opcode, hottest instructions.  usage: ncu_src.py src.csv [topN]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter()
byop = collections.Counter()
execop = collections.Counter()
inst = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or not r[ix["# Samples"]].strip().isdigit():
        continue
    s = int(r[ix['# Samples']] or 0)
    src = r[ix['Source']]
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LD', 'ST', 'RED', 'ATOM')) and '.' in op else '')
    byop[op] += s
    execop[op] += int(r[ix['Instructions Executed']] or 0)
    for c in stall_cols:
        tot[c] += int(r[ix[c]] or 0)
    inst.append((s, r[ix['Address']], src[:90], r[ix['Instructions Executed']], r[ix['Avg. Threads Executed']]))
n = sum(byop.values())
print('total samples', n)
print('stalls:', ', '.join('%s %.1f%%' % (k[6:], 100. * v / max(1, n)) for k, v in tot.most_common(9)))
print('samples by opcode:', ', '.join('%s %.1f%%' % (k, 100. * v / max(1, n)) for k, v in byop.most_common(14)))
ne = sum(execop.values())
print('executed by opcode:', ', '.join('%s %.1f%%' % (k, 100. * v / max(1, ne)) for k, v in execop.most_common(14)), ' total warp-inst', ne)
for s, a, src, ex, thr in sorted(inst, reverse=True)[:top]:
    print('%6d %5.1f%% %s  exec %s thr %s  %s' % (s, 100. * s / max(1, n), a[-5:], ex, thr, src))
