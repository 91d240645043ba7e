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
ap.add_argument('--natoms', type=int, default=None, help='first N atoms of the xyz file only')
ap.add_argument('--add-g', action='store_true',
                help='add the cc-pVQZ/def2-QZVPP g shell to C, N, O and its f shell to H (g-class stress: '
                     'stand-in for def2-QZVPP, which is not embedded)')
ap.add_argument('--ndm', type=int, default=1, help='density matrices per build (2 = UHF-like)')
args = ap.parse_args()
if args.add_g:
    from gpu4pyscf_b200.gto import basis as _b
    from gpu4pyscf_b200.gto.mole import M, read_xyz
    extra = {'C': [4, [1.011, 1.]], 'N': [4, [1.427, 1.]], 'O': [4, [1.846, 1.]], 'H': [3, [1.397, 1.]]}
    xyz = read_xyz(os.path.join(bench.MOLDIR, args.mol + '.xyz'))
    if args.natoms:
        xyz = '\n'.join(xyz.splitlines()[:args.natoms])
    els = sorted(set(l.split()[0] for l in xyz.splitlines() if l.strip()))
    mol = M(atom=xyz, basis={el: _b.load(args.basis, el) + [extra[el]] for el in els})
else:
    mol = bench.load_mol(args.mol, args.basis, args.natoms)
dm = bench.decaying_density(mol)
if args.ndm > 1:
    dm = np.stack([bench.decaying_density(mol, seed=i) for i in range(args.ndm)])
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
# B200_LAUNCH_LOG=1: the launcher prints one "JKLAUNCH ..." line per class launch to fd 2; capture them in
# launch order (class_timing keeps insertion order, one key per launch)
logf = None
if os.environ.get('B200_LAUNCH_LOG') == '1':
    sys.stderr.flush()
    logf = os.path.join(ROOT, 'gpurun_out', 'launch_log.tmp')
    os.makedirs(os.path.dirname(logf), exist_ok=True)
    saved = os.dup(2)
    fd = os.open(logf, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 2)
jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
lib.B200_set_counter(None)
shapes = {}
if logf:
    os.dup2(saved, 2)
    os.close(fd)
    lines = [l.split() for l in open(logf) if l.startswith('JKLAUNCH')]
    if len(lines) == len(opt.class_timing):
        for k, l in zip(opt.class_timing, lines):
            d = dict(zip(l[5::2], l[6::2]))
            shapes[k] = '%-6s thr %3s smem %6s b/sm %s regs %3s stack %4s grid %4s' % (
                l[1], d['threads'], d['smem'], d['blocks_per_sm'], d['regs'], d['stack'], d['grid'])
    else:
        print('launch log: %d lines for %d launches' % (len(lines), len(opt.class_timing)))
rows = []
for k, v in opt.class_timing.items():
    f = flops.f_quartet(k[:8], args.ndm) * v[3]
    rows.append((k, v[0], v[1], v[2], v[3], f))
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows); ftot = sum(r[5] for r in rows)
print('serial per-launch total %.1f ms, F_alg %.4e, %.2f TF/s' % (tot, ftot, ftot / tot / 1e9))
unc = sum(r[1] for r in rows if all(p == 1 for p in r[0][4:8]))
print('uncontracted launches: %.1f ms (%.1f%%)' % (unc, 100 * unc / tot))
for r in rows[:args.top]:
    k = r[0]
    print('(%d%d|%d%d) prim %d %d %d %d %-4s %8.2f ms %5.2f%%  npx %6d npy %6d quartets %.3e  %5.2f TF/s  %s'
          % (*k[:8], k[8] if len(k) > 8 else '', r[1], 100 * r[1] / tot, r[2], r[3], r[4], r[5] / r[1] / 1e9 if r[1] else 0,
             shapes.get(k, '')))
if args.out:
    json.dump([dict(key=list(r[0]), ms=r[1], npx=r[2], npy=r[3], quartets=r[4], flops=r[5], shape=shapes.get(r[0], '')) for r in rows], open(args.out, 'w'))
