#!/usr/bin/env python
"""Sum the per-launch ncu metrics of one J/K build (tools/one_build.py) into the counters bench.py
reports: F_hw = 2*DFMA + DMUL + DADD thread instructions, DRAM bytes, and the per-kernel-family
shares of the build time.

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/build_counters.csv \\
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\\
smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__inst_executed.sum,\\
sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active python tools/one_build.py
    python tools/ncu_build_counters.py gpurun_out/build_counters.csv profiles/ncu_r2/build_counters_113_Taxol_def2-tzvp.json
"""
import csv, sys, json, re, collections

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()          # launch id -> {metric: value}
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    key = r[ix['ID']]
    d = per.setdefault(key, {'kernel': r[ix['Kernel Name']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    unit = r[ix['Metric Unit']]
    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1., 'usecond': 1e-3, 'msecond': 1., 'second': 1e3,
             'nsecond': 1e-6, 'us': 1e-3, 'ms': 1., 'ns': 1e-6, 's': 1e3}.get(unit, 1.)
    d[r[ix['Metric Name']]] = v * scale
fam = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for d in per.values():
    m = re.search(r'(rys_jk_\w+|condense_kernel|\w+_kernel)', d['kernel'])
    name = m.group(1) if m else d['kernel'][:40]
    ms = d.get('gpu__time_duration.sum', 0.)
    dfma = d.get('smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 0.)
    dmul = d.get('smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 0.)
    dadd = d.get('smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 0.)
    vals = dict(launches=1, ms=ms, f_hw=2 * dfma + dmul + dadd, dfma=dfma, dmul=dmul, dadd=dadd,
                dram_bytes=d.get('dram__bytes_read.sum', 0.) + d.get('dram__bytes_write.sum', 0.),
                warp_inst=d.get('smsp__inst_executed.sum', 0.),
                fp64_pipe_ms=ms * d.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 0.) / 100.)
    fam[name].update(vals)
    tot.update(vals)
out = dict(tot)
out['fp64_pipe_active_pct_time_weighted'] = 100. * tot['fp64_pipe_ms'] / tot['ms'] if tot['ms'] else None
out['families'] = {k: dict(v) for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms'])}
out['source'] = ('ncu --profile-from-start off --clock-control none, one build of tools/one_build.py, summed over %d '
                 'launches (times are serialised, cold-cache ncu times: use shares, not absolutes)' % tot['launches'])
json.dump(out, open(sys.argv[2], 'w'), indent=1)
print('launches %d  ms %.1f  F_hw %.4e  DRAM %.3e B  fp64 pipe %.1f%%' % (
    tot['launches'], tot['ms'], tot['f_hw'], tot['dram_bytes'], out['fp64_pipe_active_pct_time_weighted']))
for k, v in out['families'].items():
    print('  %-28s launches %4d  ms %8.1f (%4.1f%%)  F_hw %.3e  DRAM %.3e' % (
        k, v['launches'], v['ms'], 100 * v['ms'] / tot['ms'], v['f_hw'], v['dram_bytes']))
