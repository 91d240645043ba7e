#!/usr/bin/env python
"""Shared-memory wavefronts per instruction from an `ncu --page source --csv` dump: the top
consumers, with executed count, wavefronts per execution and the excess over the ideal.
usage: ncu_smem.py src.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
seen = set()
out = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix['Instructions Executed']].strip().isdigit():
        continue
    if r[ix['Address']] in seen:
        continue
    seen.add(r[ix['Address']])
    w = int(r[ix['L1 Wavefronts Shared']] or 0)
    if w:
        ex = int(r[ix['Instructions Executed']])
        out.append((w, ex, int(r[ix['L1 Wavefronts Shared Ideal']] or 0), r[ix['Address']][-5:], r[ix['Source']].strip()[:70],
                    r[ix['Avg. Threads Executed']]))
tot = sum(o[0] for o in out)
print('total shared wavefronts', tot, ' ideal', sum(o[2] for o in out))
import collections
byop = collections.Counter()
for o in out:
    byop[o[4].split()[0] if not o[4].startswith('@') else o[4].split()[1]] += o[0]
print(' '.join('%s:%.1f%%' % (k, 100. * v / tot) for k, v in byop.most_common(6)))
for w, ex, ideal, addr, src, thr in sorted(out, reverse=True)[:top]:
    print('%5.1f%%  %5.2f wf/exec (ideal %5.2f)  exec %9d thr %s  %s  %s' % (100. * w / tot, w / ex, ideal / ex, ex, thr, addr, src))
