#!/usr/bin/env python
"""Split an `ncu --page source --csv` dump into regions of equal execution count (a proxy for
the phases of a kernel) and print, per region: static instructions, executed warp-instructions,
FP64 share and the opcode mix.  usage: ncu_regions.py src.csv [min_share_percent]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
regions = []
cur = None
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix['Instructions Executed']].strip().isdigit():
        continue
    ex = int(r[ix['Instructions Executed']])
    toks = r[ix['Source']].split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '')
    op = op.split('.')[0]
    thr = float(r[ix['Avg. Threads Executed']] or 0)
    if cur is None or cur['ex'] != ex:
        cur = dict(ex=ex, n=0, ops=collections.Counter(), addr=r[ix['Address']][-5:], thr=0.)
        regions.append(cur)
    cur['n'] += 1
    cur['ops'][op] += 1
    cur['thr'] += thr
tot = sum(g['ex'] * g['n'] for g in regions)
print('total executed warp-instructions', tot)
# merge regions with the same execution count
merged = collections.OrderedDict()
for g in regions:
    m = merged.setdefault(g['ex'], dict(ex=g['ex'], n=0, ops=collections.Counter(), thr=0., first=g['addr']))
    m['n'] += g['n']; m['ops'] += g['ops']; m['thr'] += g['thr']
fp = ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU')
for m in sorted(merged.values(), key=lambda m: -m['ex'] * m['n']):
    share = 100. * m['ex'] * m['n'] / tot
    if share < minshare:
        continue
    nfp = sum(v for k, v in m['ops'].items() if k in fp)
    print('exec %10d x %5d instr = %5.1f%%  fp64 %4.1f%%  thr %4.1f  first %s  %s' % (
        m['ex'], m['n'], share, 100. * nfp / m['n'], m['thr'] / m['n'], m['first'],
        ' '.join('%s:%d' % kv for kv in m['ops'].most_common(9))))
