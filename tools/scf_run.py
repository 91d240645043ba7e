#!/usr/bin/env python
"""RHF to convergence on the GPU J/K build (scf/hf.py, incremental Fock builds), one-electron integrals
from the CPU oracle -- BASELINE configs 2/3 with an SCF density instead of the synthetic bench density.
Prints per cycle: energy, J/K (get_veff) time; at the end the time and the surviving-quartet count of one
full J/K build at the final density next to the same numbers for the bench density.
usage: python tools/scf_run.py [--mol 113_Taxol] [--basis def2-tzvp] [--max-cycle 40]"""
import os, sys, time, ctypes, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from gpu4pyscf_b200.scf import jk, hf
from gpu4pyscf_b200.lib import loader
from oracle import oracle

ap = argparse.ArgumentParser()
ap.add_argument('--mol', default='113_Taxol')
ap.add_argument('--basis', default='def2-tzvp')
ap.add_argument('--max-cycle', type=int, default=40)
ap.add_argument('--conv-tol', type=float, default=1e-8)
args = ap.parse_args()
mol = bench.load_mol(args.mol, args.basis)
t0 = time.time()
s, t, v = oracle.int1e(mol)
print('one-electron integrals (CPU oracle): %.1f s, nao %d' % (time.time() - t0, mol.nao), flush=True)
mf = hf.RHF(mol)
mf.hcore, mf.ovlp = t + v, s
mf.max_cycle = args.max_cycle
mf.conv_tol = args.conv_tol
veff_ms = []
orig = mf.get_veff


def timed_veff(*a, **k):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(*a, **k)
    e1.record(); e1.synchronize()
    veff_ms.append(e0.elapsed_time(e1))
    print('  get_veff %d: %.0f ms (one J/K pass, %s density)' % (len(veff_ms), veff_ms[-1],
          'full' if len(a) < 3 and 'dm_last' not in k else 'difference'), flush=True)
    return out


mf.get_veff = timed_veff
t0 = time.time()
e = mf.kernel()
print('RHF %s/%s: E = %.10f  converged %s  cycles %d  wall %.1f s' % (args.mol, args.basis, e, mf.converged, mf.cycles,
                                                                     time.time() - t0), flush=True)
lib = loader.load_library()
opt = mf._vhfopt(mol.omega)


def build_stats(dm, label):
    dms = opt.apply_coeff_C_mat_CT(dm.reshape(1, mol.nao, mol.nao)).contiguous()
    opt.get_jk(dms, 1)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); opt.get_jk(dms, 1); e1.record(); e1.synchronize()
    c = torch.zeros(1, dtype=torch.int64, device='cuda')
    lib.B200_set_counter(ctypes.c_void_p(c.data_ptr()))
    opt.get_jk(dms, 1)
    torch.cuda.synchronize()
    lib.B200_set_counter(None)
    print('J/K build, %s: %.1f ms, %.4e quartets' % (label, e0.elapsed_time(e1), float(c.item())), flush=True)


build_stats(mf.make_rdm1(), 'SCF density (%s)' % ('converged' if mf.converged else 'after %d cycles' % mf.cycles))
build_stats(torch.as_tensor(bench.decaying_density(mol), device='cuda'), 'synthetic bench density')
