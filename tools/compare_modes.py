#!/usr/bin/env python
"""Row-stream launches vs flat launches on the same input: max|dJ|, max|dK| and build times."""
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
res = {}
for rows in ('0', '1'):
    os.environ['B200_ROWS'] = rows
    opt = jk._VHFOpt(mol).build()
    tasks = opt.class_tasks()
    for _ in range(2):
        vj, vk = jk.get_jk(mol, dm, hermi=1, vhfopt=opt)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); vj, vk = jk.get_jk(mol, dm, hermi=1, vhfopt=opt); e1.record(); e1.synchronize()
    print('rows=%s: %d launches (%d row-stream), get_jk %.1f ms' %
          (rows, len(tasks), sum(t[0] == 'row' for t in tasks), e0.elapsed_time(e1)), flush=True)
    res[rows] = (vj, vk)
print('max|dJ| %.3e  max|dK| %.3e   (max|J| %.3e max|K| %.3e)' % (
    (res['0'][0] - res['1'][0]).abs().max(), (res['0'][1] - res['1'][1]).abs().max(),
    res['0'][0].abs().max(), res['0'][1].abs().max()))
