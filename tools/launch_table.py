#!/usr/bin/env python
"""Per-launch table of one J/K build: (class, nprim) -> ms, X/Y pair counts, surviving quartets, F_alg, TFLOP/s.
Usage: python tools/launch_table.py [--mol 113_Taxol] [--basis def2-tzvp] [--out gpurun_out/launch_table.json]"""
import os, sys, json, argparse, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from gpu4pyscf_b200.scf import jk, flops
from gpu4pyscf_b200.lib import loader

ap = argparse.ArgumentParser()
ap.add_argument('--mol', default='113_Taxol')
ap.add_argument('--basis', default='def2-tzvp')
ap.add_argument('--out', default=None)
ap.add_argument('--top', type=int, default=60)
args = ap.parse_args()
mol = bench.load_mol(args.mol, args.basis)
dm = bench.decaying_density(mol)
opt = jk._VHFOpt(mol).build()
lib = loader.load_library()
dmt = torch.as_tensor(dm, device='cuda')
for _ in range(2):
    jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); jk.get_jk(mol, dmt, hermi=1, vhfopt=opt); e1.record(); e1.synchronize()
print('get_jk %.1f ms' % e0.elapsed_time(e1))
opt.counter = torch.zeros(1, dtype=torch.int64, device='cuda')
lib.B200_set_counter(ctypes.c_void_p(opt.counter.data_ptr()))
opt.class_timing = {}
jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
lib.B200_set_counter(None)
rows = []
for k, v in opt.class_timing.items():
    f = flops.f_quartet(k[:8]) * v[3]
    rows.append((k, v[0], v[1], v[2], v[3], f))
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows); ftot = sum(r[5] for r in rows)
print('serial per-launch total %.1f ms, F_alg %.4e, %.2f TF/s' % (tot, ftot, ftot / tot / 1e9))
unc = sum(r[1] for r in rows if all(p == 1 for p in r[0][4:8]))
print('uncontracted launches: %.1f ms (%.1f%%)' % (unc, 100 * unc / tot))
for r in rows[:args.top]:
    k = r[0]
    print('(%d%d|%d%d) prim %d %d %d %d %-4s %8.2f ms %5.2f%%  npx %6d npy %6d quartets %.3e  %.2f TF/s'
          % (*k[:8], k[8] if len(k) > 8 else '', r[1], 100 * r[1] / tot, r[2], r[3], r[4], r[5] / r[1] / 1e9 if r[1] else 0))
if args.out:
    json.dump([dict(key=list(r[0]), ms=r[1], npx=r[2], npy=r[3], quartets=r[4], flops=r[5]) for r in rows], open(args.out, 'w'))
