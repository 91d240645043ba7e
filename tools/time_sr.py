#!/usr/bin/env python
"""Taxol/def2-TZVP: K-only builds for 1/r, erf(w r)/r and erfc(w r)/r (one pass, two operator terms per
quartet, operator-specific Schwarz bounds) against the two-pass composition full - long range."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gpu4pyscf_b200.scf import jk

mol = bench.load_mol('113_Taxol', 'def2-tzvp')
dm = torch.as_tensor(bench.decaying_density(mol), device='cuda')
opt = jk._VHFOpt(mol).build()
dms = opt.apply_coeff_C_mat_CT(dm.reshape(1, mol.nao, mol.nao)).contiguous()


def timed(fn, n=2):
    fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n, out


import ctypes
from gpu4pyscf_b200.lib import loader
lib = loader.load_library()


def count(fn):
    c = torch.zeros(1, dtype=torch.int64, device='cuda')
    lib.B200_set_counter(ctypes.c_void_p(c.data_ptr()))
    fn()
    torch.cuda.synchronize()
    lib.B200_set_counter(None)
    return int(c.item())


w = 0.3
t_full, k_full = timed(lambda: opt.get_k(dms, 1))
t_lr, k_lr = timed(lambda: opt.get_k(dms, 1, omega=w))
os.environ['B200_SR_FILTER'] = '0'
t_sr0, k_sr0 = timed(lambda: opt.get_k(dms, 1, omega=-w))
n_sr0 = count(lambda: opt.get_k(dms, 1, omega=-w))
os.environ['B200_SR_FILTER'] = '1'
t_sr, k_sr = timed(lambda: opt.get_k(dms, 1, omega=-w))
n_sr = count(lambda: opt.get_k(dms, 1, omega=-w))
n_full = count(lambda: opt.get_k(dms, 1))
npairs = {k: sum(len(v[0]) for v in pl.cache.values()) for k, pl in opt._lists.items()}
print('K 1/r        %.1f ms   pairs %d' % (t_full, npairs[((1., 0.),)]))
print('K erf(%.1fr)/r  %.1f ms   pairs %d' % (w, t_lr, npairs[((1., w),)]))
print('K erfc(%.1fr)/r %.1f ms   pairs %d   one pass, SR Schwarz bounds only: %d quartets (full-range K: %d); '
      'two-pass composition would cost %.1f ms' % (w, t_sr0, npairs[((1., 0.), (-1., w))], n_sr0, n_full, t_full + t_lr))
print('K erfc(%.1fr)/r %.1f ms   with the short-range quartet filter (s_estimator): %d quartets; '
      'max|K_filter - K_nofilter| = %.2e' % (w, t_sr, n_sr, float((k_sr - k_sr0).abs().max())))
print('max|K_sr - (K_full - K_lr)| = %.2e  (max|K_sr| %.2e)' % (float((k_sr - (k_full - k_lr)).abs().max()), float(k_sr.abs().max())))
