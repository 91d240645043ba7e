#!/usr/bin/env python
"""Time one J/K build (and optionally every angular class) on the GPU."""
import os, sys, time, json, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from gpu4pyscf_b200.gto.mole import M, read_xyz
from gpu4pyscf_b200.scf import jk

ap = argparse.ArgumentParser()
ap.add_argument('--mol', default='020_Vitamin_C')
ap.add_argument('--basis', default='def2-tzvp')
ap.add_argument('--reps', type=int, default=2)
ap.add_argument('--classes', action='store_true')
ap.add_argument('--check', action='store_true')
ap.add_argument('--out', default=None)
args = ap.parse_args()

xyz = os.path.join(ROOT, 'gpu4pyscf_b200', 'data', 'molecules', args.mol + '.xyz')
mol = M(atom=read_xyz(xyz), basis=args.basis)
nao = mol.nao
print('mol', args.mol, 'nao', nao, 'nbas', mol.nbas, flush=True)

def decaying_density(mol, seed=0, beta=0.8):
    rng = np.random.RandomState(seed)
    ao_atom = np.repeat(mol._bas[:, 0], np.diff(mol.ao_loc_nr()))
    R = mol.atom_coords()
    dist = np.linalg.norm(R[ao_atom][:, None] - R[ao_atom][None], axis=-1)
    d = (rng.rand(nao, nao) - .5) * np.exp(-beta * dist)
    d = d + d.T + np.eye(nao)
    return d

dm = decaying_density(mol)
t0 = time.time()
opt = jk._VHFOpt(mol).build()
torch.cuda.synchronize()
print('build %.3f s' % (time.time() - t0), 'npairs', sum(len(v[0]) for v in opt.bas_pair_cache.values()),
      'groups', [tuple(x) for x in opt.sorted_mol.uniq_l_ctr], list(opt.sorted_mol.l_ctr_counts), flush=True)
dmt = torch.as_tensor(dm, device='cuda')
res = {}
for rep in range(args.reps + 1):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    vj, vk = jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
    e1.record(); e1.synchronize()
    print('get_jk rep %d: %.3f ms' % (rep, e0.elapsed_time(e1)), flush=True)
    res.setdefault('ms', []).append(e0.elapsed_time(e1))
if args.classes:
    opt.class_timing = {}
    vj, vk = jk.get_jk(mol, dmt, hermi=1, vhfopt=opt)
    rows = sorted(opt.class_timing.items(), key=lambda kv: -kv[1][0])
    tot = sum(v[0] for _, v in rows)
    print('per-class total %.3f ms' % tot)
    for k, v in rows[:40]:
        print('  (%d%d|%d%d) prim %d%d%d%d  %9.3f ms  %5.1f%%  npx %d npy %d' % (*k, v[0], 100 * v[0] / tot, v[1], v[2]))
    res['classes'] = {str(k): v for k, v in rows}
    opt.class_timing = None
if args.check:
    from oracle import oracle
    t0 = time.time()
    rj, rk = oracle.get_jk(mol, dm, hermi=1, screen=True, tol=1e-14)
    print('oracle %.1f s' % (time.time() - t0))
    print('max|dJ| %.3e max|dK| %.3e  (max|J| %.3e)' % (abs(vj.cpu().numpy() - rj).max(), abs(vk.cpu().numpy() - rk).max(), abs(rj).max()))
if args.out:
    json.dump(res, open(args.out, 'w'))
