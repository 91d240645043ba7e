#!/usr/bin/env python
"""One J/K build of the bench workload between cudaProfilerStart/Stop (after a warm-up build), for
`ncu --profile-from-start off`.  usage: python tools/one_build.py [--mol 113_Taxol] [--basis def2-tzvp]"""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gpu4pyscf_b200.scf import jk

ap = argparse.ArgumentParser()
ap.add_argument('--mol', default='113_Taxol')
ap.add_argument('--basis', default='def2-tzvp')
args = ap.parse_args()
mol = bench.load_mol(args.mol, args.basis)
dm = torch.as_tensor(bench.decaying_density(mol), device='cuda')
opt = jk._VHFOpt(mol).build()
opt.n_streams = 1
dms = opt.apply_coeff_C_mat_CT(dm.reshape(1, mol.nao, mol.nao)).contiguous()
opt.get_jk(dms, 1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
opt.get_jk(dms, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
