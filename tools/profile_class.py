#!/usr/bin/env python
"""Per-class timing / flop-rate table for one J/K build (optionally restricted
to a few classes so the run can be wrapped in ncu)."""
import os, sys, time, json, argparse, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from gpu4pyscf_b200.gto.mole import M, read_xyz
from gpu4pyscf_b200.scf import jk, flops
from gpu4pyscf_b200.lib import loader

ap = argparse.ArgumentParser()
ap.add_argument('--mol', default='113_Taxol')
ap.add_argument('--basis', default='def2-tzvp')
ap.add_argument('--classes', default='')       # "2,2,2,2;2,1,2,1"
ap.add_argument('--top', type=int, default=30)
ap.add_argument('--out', default=None)
ap.add_argument('--beta', type=float, default=0.8)
ap.add_argument('--detail', default=None, help='json: one row per (l, nprim) launch')
args = ap.parse_args()
xyz = os.path.join(ROOT, 'gpu4pyscf_b200', 'data', 'molecules', args.mol + '.xyz')
mol = M(atom=read_xyz(xyz), basis=args.basis)
nao = mol.nao
rng = np.random.RandomState(0)
ao_atom = np.repeat(mol._bas[:, 0], np.diff(mol.ao_loc_nr()))
R = mol.atom_coords()
dist = np.linalg.norm(R[ao_atom][:, None] - R[ao_atom][None], axis=-1)
dm = (rng.rand(nao, nao) - .5) * np.exp(-args.beta * dist)
dm = dm + dm.T + np.eye(nao)
opt = jk._VHFOpt(mol).build()
lib = loader.load_library()
opt.counter = torch.zeros(1, dtype=torch.int64, device='cuda')
lib.B200_set_counter(ctypes.c_void_p(opt.counter.data_ptr()))
if args.classes:
    opt.only_classes = set(tuple(int(v) for v in c.split(',')) for c in args.classes.split(';'))
dmt = torch.as_tensor(dm, device='cuda')
jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)      # warm-up
opt.class_timing = {}
torch.cuda.synchronize()
t0 = time.time()
jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
torch.cuda.synchronize()
wall = time.time() - t0
agg = {}
for k, v in opt.class_timing.items():
    a = agg.setdefault(k[:4], [0., 0, 0.])
    a[0] += v[0]; a[1] += v[3]; a[2] += v[3] * flops.f_quartet(k)
tot_ms = sum(a[0] for a in agg.values()); tot_fl = sum(a[2] for a in agg.values()); tot_q = sum(a[1] for a in agg.values())
print('variant', os.environ.get('GPU4PYSCF_B200_VARIANT'), 'total %.1f ms (wall %.1f)  quartets %.3e  F_alg %.3e  -> %.3f TFLOP/s' % (tot_ms, wall * 1e3, tot_q, tot_fl, tot_fl / tot_ms / 1e9))
rows = sorted(agg.items(), key=lambda kv: -kv[1][0])
for k, a in rows[:args.top]:
    print('  (%d%d|%d%d) %9.2f ms %5.1f%%  quartets %.3e  %7.3f TF/s' % (*k, a[0], 100 * a[0] / tot_ms, a[1], a[2] / max(a[0], 1e-9) / 1e9))
if args.detail:
    json.dump([{'lprim': list(k), 'ms': v[0], 'npx': v[1], 'npy': v[2], 'quartets': v[3], 'flops': v[3] * flops.f_quartet(k)}
               for k, v in opt.class_timing.items()], open(args.detail, 'w'))
if args.out:
    json.dump({'total_ms': tot_ms, 'quartets': tot_q, 'f_alg': tot_fl,
               'classes': {str(k): a for k, a in rows}}, open(args.out, 'w'))
